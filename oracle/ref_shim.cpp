// TEST INFRASTRUCTURE ONLY (never linked or loaded by the product path).
//
// extern "C" wrapper (our own code) around the UNMODIFIED reference CPU classes so that python/ctypes can
// (a) pin oracle/vlq_oracle.c against the reference itself, (b) generate tests/golden/*.npz, and (c) time the
// reference's CPU implementation in bench.py (--impl reference / cpu_baseline.kind == "reference").
// It is compiled together with the reference sources where they lie (see oracle/Makefile, target `ref`).
//
// Reference entry points used: IndexFlatL2::add/search (IndexFlat.cpp:36-56), Clustering::train
// (Clustering.cpp:66-206), ProductQuantizer::train/compute_codes (ProductQuantizer.cpp:236-308,385-407),
// IndexIVFPQ::add/search/precompute_table (IndexIVFPQ.cpp:192-272,392-459,1063-1081),
// IndexShards::search (MetaIndexes.cpp:486-563), float_maxheap_array_t (Heap.h).
#include <cstdio>
#include <cstring>
#include <vector>
#include <omp.h>

#include "Clustering.h"
#include "Heap.h"
#include "IndexFlat.h"
#include "IndexIVFPQ.h"
#include "IndexPQ.h"
#include "MetaIndexes.h"
#include "ProductQuantizer.h"
#include "utils.h"

using namespace faiss;

extern "C" {

int ref_num_threads() { return omp_get_max_threads(); }

// nearest-k base vectors by IndexFlatL2 (knn_L2sqr: BLAS blocks + max-heap when nq >= 20, SSE otherwise)
void ref_flat_search(int d, long nb, const float* xb, long nq, const float* xq, long k, float* D, long* I) {
  IndexFlatL2 index(d);
  index.add(nb, xb);
  index.search(nq, xq, k, D, I);
}

// Clustering::train over an IndexFlatL2 assigner (the CPU twin of the loop the GPU flat index serves)
void ref_kmeans(int d, int k, long n, const float* x, int niter, long seed, float* centroids_out) {
  ClusteringParameters cp;
  cp.niter = niter;
  cp.seed = (int)seed;
  Clustering clus(d, k, cp);
  IndexFlatL2 index(d);
  clus.train(n, x, index);
  memcpy(centroids_out, clus.centroids.data(), sizeof(float) * (size_t)d * k);
}

void ref_rand_perm(int* perm, long n, long seed) { rand_perm(perm, n, seed); }

int ref_km_update_centroids(const float* x, float* centroids, long* assign, long d, long k, long n) {
  return km_update_centroids(x, centroids, assign, d, k, n);
}

void ref_pq_train(int d, int M, int nbits, long n, const float* x, float* centroids_out) {
  ProductQuantizer pq(d, M, nbits);
  pq.train((int)n, x);
  memcpy(centroids_out, pq.centroids.data(), sizeof(float) * pq.centroids.size());
}

void ref_pq_compute_codes(int d, int M, int nbits, const float* centroids, long n, const float* x,
                          unsigned char* codes) {
  ProductQuantizer pq(d, M, nbits);
  memcpy(pq.centroids.data(), centroids, sizeof(float) * pq.centroids.size());
  pq.compute_codes(x, codes, n);
}

// exact top-k ascending of each row of vals[n][m] through the reference heap (Heap.h:89-143,296-323)
void ref_heap_topk(long n, long m, long k, const float* vals, float* D, long* I) {
  float_maxheap_array_t res = {size_t(n), size_t(k), I, D};
  res.heapify();
  for (long i = 0; i < n; i++) {
    float* simi = res.get_val(i);
    long* idxi = res.get_ids(i);
    for (long j = 0; j < m; j++) {
      float dis = vals[i * m + j];
      if (dis < simi[0]) {
        maxheap_pop(k, simi, idxi);
        maxheap_push(k, simi, idxi, dis, j);
      }
    }
  }
  res.reorder();
}

// ---- IndexIVFPQ with externally supplied codebooks (so oracle / CUDA path can be compared on equal terms) ----
struct RefIVFPQ {
  IndexFlatL2* quantizer;
  IndexIVFPQ* index;
};

void* ref_ivfpq_new(int d, long nlist, int M, int nbits, const float* coarse, const float* pq_centroids,
                    int use_precomputed_table) {
  RefIVFPQ* h = new RefIVFPQ;
  h->quantizer = new IndexFlatL2(d);
  h->index = new IndexIVFPQ(h->quantizer, d, nlist, M, nbits);
  h->index->use_precomputed_table = use_precomputed_table;
  if (coarse) {
    h->quantizer->add(nlist, coarse);
    memcpy(h->index->pq.centroids.data(), pq_centroids, sizeof(float) * h->index->pq.centroids.size());
    h->index->is_trained = true;
    h->index->quantizer_trains_alone = false;
    if (use_precomputed_table) h->index->precompute_table();
  }
  return h;
}

void ref_ivfpq_train(void* hv, long n, const float* x) { ((RefIVFPQ*)hv)->index->train(n, x); }

void ref_ivfpq_get_codebooks(void* hv, float* coarse, float* pq_centroids) {
  RefIVFPQ* h = (RefIVFPQ*)hv;
  memcpy(coarse, h->quantizer->xb.data(), sizeof(float) * h->quantizer->xb.size());
  memcpy(pq_centroids, h->index->pq.centroids.data(), sizeof(float) * h->index->pq.centroids.size());
}

void ref_ivfpq_add(void* hv, long n, const float* x) { ((RefIVFPQ*)hv)->index->add(n, x); }

void ref_ivfpq_search(void* hv, long nq, const float* xq, long k, long nprobe, float* D, long* I) {
  RefIVFPQ* h = (RefIVFPQ*)hv;
  h->index->nprobe = nprobe;
  h->index->search(nq, xq, k, D, I);
}

long ref_ivfpq_list_size(void* hv, long list) { return (long)((RefIVFPQ*)hv)->index->ids[list].size(); }

void ref_ivfpq_get_list(void* hv, long list, long* ids, unsigned char* codes) {
  RefIVFPQ* h = (RefIVFPQ*)hv;
  memcpy(ids, h->index->ids[list].data(), sizeof(long) * h->index->ids[list].size());
  memcpy(codes, h->index->codes[list].data(), h->index->codes[list].size());
}

void ref_ivfpq_free(void* hv) {
  RefIVFPQ* h = (RefIVFPQ*)hv;
  delete h->index;
  delete h->quantizer;
  delete h;
}

// ---- IndexShards over flat shards: defines the k-way merge semantics (MetaIndexes.cpp:290-347,486-563) ----
void ref_shards_flat_search(int d, int nshard, const long* shard_sizes, const float* xb, long nq, const float* xq,
                            long k, float* D, long* I) {
  IndexShards shards(d, /*threaded=*/false, /*successive_ids=*/true);
  std::vector<IndexFlatL2*> subs;
  long off = 0;
  for (int s = 0; s < nshard; s++) {
    IndexFlatL2* f = new IndexFlatL2(d);
    f->add(shard_sizes[s], xb + (size_t)off * d);
    off += shard_sizes[s];
    subs.push_back(f);
    shards.add_shard(f);
  }
  shards.search(nq, xq, k, D, I);
  for (auto* f : subs) delete f;
}


// MultiIndexQuantizer::search (IndexPQ.cpp:813-855) on given sub-space codebooks (M, 2^nbits, d/M): the IMI coarse
// quantizer of the sift1b_imi_pq baselines (next row f3)
void ref_imi_search(int d, int M, int nbits, const float* centroids, long n, const float* x, long k, float* D, long* I) {
  MultiIndexQuantizer miq(d, M, nbits);
  memcpy(miq.pq.centroids.data(), centroids, sizeof(float) * miq.pq.centroids.size());
  miq.is_trained = true;
  miq.ntotal = 1;
  for (int m = 0; m < M; m++) miq.ntotal *= miq.pq.ksub;
  miq.search(n, x, k, D, I);
}

// ---- the IMI-PQ baseline index of tests/sift1b_imi_pq.cpp:216-236 (BASELINE configs[4]): MultiIndexQuantizer(d, 2, nbits_coarse)
// as coarse quantizer of an IndexIVFPQ with 2^(2*nbits_coarse) lists, quantizer_trains_alone = true
struct RefIMIPQ {
  MultiIndexQuantizer* miq;
  IndexIVFPQ* index;
};
void* ref_imipq_new(int d, int nbits_coarse, int M, int nbits) {
  RefIMIPQ* h = new RefIMIPQ;
  h->miq = new MultiIndexQuantizer(d, 2, nbits_coarse);
  h->index = new IndexIVFPQ(h->miq, d, (size_t)1 << (2 * nbits_coarse), M, nbits);
  h->index->quantizer_trains_alone = true;
  return h;
}
void ref_imipq_train(void* hv, long n, const float* x) { ((RefIMIPQ*)hv)->index->train(n, x); }
void ref_imipq_add(void* hv, long n, const float* x) { ((RefIMIPQ*)hv)->index->add(n, x); }
void ref_imipq_search(void* hv, long nq, const float* xq, long k, long nprobe, float* D, long* I) {
  RefIMIPQ* h = (RefIMIPQ*)hv;
  h->index->nprobe = nprobe;
  h->index->search(nq, xq, k, D, I);
}
// trained codebooks: coarse = (2, 2^nbits_coarse, d/2), pq = (M, 2^nbits, d/M)
void ref_imipq_get_codebooks(void* hv, float* coarse, float* pq) {
  RefIMIPQ* h = (RefIMIPQ*)hv;
  memcpy(coarse, h->miq->pq.centroids.data(), sizeof(float) * h->miq->pq.centroids.size());
  memcpy(pq, h->index->pq.centroids.data(), sizeof(float) * h->index->pq.centroids.size());
}
void ref_imipq_free(void* hv) {
  RefIMIPQ* h = (RefIMIPQ*)hv;
  delete h->index;
  delete h->miq;
  delete h;
}
}  // extern "C"
