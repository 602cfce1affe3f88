// SQUAREM EM over counts_umi on the GPU (src/em.rs:48-138 EmProblem for EqClassCounts, :232-311 squarem, :315-336
// diffs), f64 throughout, plus the genotype caller (src/em.rs:361-434) on the host.
//
// The EM step is a sparse gather/scatter: per class norm = sum(theta[tx]); per tx theta2[tx] += theta[tx]/norm*count.
// The reference iterates a std HashMap (summation order unspecified, hence the 1e-6 tolerance); here the scatter is a
// deterministic gather over the transposed CSR (one warp per tx, fixed tree order), so runs are reproducible.
// No tensor cores: nothing here is a dense contraction.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <set>
#include <vector>

#include "caller.hpp"
#include "device_index.hpp"

using namespace hlac;

namespace {

__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
  return v;
}

// one warp per class: norm[c] = sum theta[tx]; optionally ll_term[c] = count * ln(norm) (em.rs:119-137)
__global__ void k_class_norm(const uint64_t* __restrict__ off, const uint32_t* __restrict__ tx, const uint32_t* __restrict__ cnt,
                             uint64_t nclasses, const double* __restrict__ theta, double* norm, double* ll_term) {
  const uint64_t c = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (c >= nclasses) return;
  double s = 0.0;
  for (uint64_t q = off[c] + lane; q < off[c + 1]; q += 32) s += theta[tx[q]];
  s = warp_sum(s);
  if (lane == 0) {
    if (norm) norm[c] = s;
    if (ll_term) ll_term[c] = off[c] == off[c + 1] ? 0.0 : (double)cnt[c] * log(s);
  }
}

// one warp per tx over the transposed CSR: raw[t] = sum_c theta[t] / norm[c] * count[c] (em.rs:84-95)
__global__ void k_tx_gather(const uint64_t* __restrict__ toff, const uint32_t* __restrict__ tcls, const uint32_t* __restrict__ cnt,
                            uint32_t nitems, const double* __restrict__ theta, const double* __restrict__ norm, double* raw) {
  const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (t >= nitems) return;
  const double th = theta[t];
  double s = 0.0;
  for (uint64_t q = toff[t] + lane; q < toff[t + 1]; q += 32) { const uint32_t c = tcls[q]; s += th / norm[c] * (double)cnt[c]; }
  s = warp_sum(s);
  if (lane == 0) raw[t] = s;
}

// single-block deterministic reductions: out[0] = sum(a), out[1] = sum(b) (b optional)
__global__ void k_sum2(const double* __restrict__ a, const double* __restrict__ b, uint64_t n, double* out) {
  __shared__ double sa[32], sb[32];
  double x = 0.0, y = 0.0;
  for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) { x += a[i]; if (b) y += b[i]; }
  x = warp_sum(x); y = warp_sum(y);
  if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = x; sb[threadIdx.x >> 5] = y; }
  __syncthreads();
  if (threadIdx.x < 32) {
    x = threadIdx.x < (blockDim.x >> 5) ? sa[threadIdx.x] : 0.0; y = threadIdx.x < (blockDim.x >> 5) ? sb[threadIdx.x] : 0.0;
    x = warp_sum(x); y = warp_sum(y);
    if (threadIdx.x == 0) { out[0] = x; out[1] = y; }
  }
}

__global__ void k_scale_by_inv(double* v, uint32_t n, const double* total, int reciprocal) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (reciprocal) v[i] *= 1.0 / total[0];   // reg(): inv_norm = 1/norm; theta *= inv_norm (em.rs:69-72)
  else v[i] = v[i] / total[0];              // f(): theta2[i] / total_counts (em.rs:102)
}

// r = t1 - t0; v = t2 - t1 - r; sq[i] = r^2, sq2[i] = v^2 (em.rs:255-261)
__global__ void k_rv(const double* t0, const double* t1, const double* t2, uint32_t n, double* r, double* v, double* rsq, double* vsq) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double ri = t1[i] - t0[i];
  const double vi = t2[i] - t1[i] - ri;
  r[i] = ri; v[i] = vi; rsq[i] = ri * ri; vsq[i] = vi * vi;
}

// theta_sq = t0 - 2 alpha r + alpha^2 v, then reg()'s clamp (em.rs:271-273, 56-67)
__global__ void k_theta_sq(const double* t0, const double* r, const double* v, uint32_t n, double alpha, double* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double asq = alpha * alpha;
  double x = t0[i] - 2.0 * alpha * r[i] + asq * v[i];
  if (x > 1.0) x = 1.0;
  if (x < 0.0) x = 1e-15;
  out[i] = x;
}

// diffs() (em.rs:315-336): out[0] = max rel (only where new > EM_CARE_TH), out[1] = max abs. NaN-propagation follows
// the reference's `if a > max` comparisons (a NaN never replaces the max).
__global__ void k_diffs(const double* t1, const double* t2, uint32_t n, double* out) {
  __shared__ double sr[32], sa[32];
  double mr = 0.0, ma = 0.0;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double a = fabs(t1[i] - t2[i]);
    const double rel = a / t1[i];
    if (a > ma) ma = a;
    if (t2[i] > EM_CARE_TH && rel > mr) mr = rel;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double x = __shfl_down_sync(0xFFFFFFFFu, mr, o), y = __shfl_down_sync(0xFFFFFFFFu, ma, o);
    if (x > mr) mr = x;
    if (y > ma) ma = y;
  }
  if ((threadIdx.x & 31) == 0) { sr[threadIdx.x >> 5] = mr; sa[threadIdx.x >> 5] = ma; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (unsigned w = 1; w < (blockDim.x >> 5); w++) { if (sr[w] > mr) mr = sr[w]; if (sa[w] > ma) ma = sa[w]; }
    out[0] = mr; out[1] = ma;
  }
}

struct EmDevice {
  cudaStream_t st;
  uint32_t nitems; uint64_t nclasses, nnz;
  DevBuf<uint64_t> off, toff; DevBuf<uint32_t> tx, cnt, tcls;
  DevBuf<double> norm, llt, raw, scal;
  uint64_t launches = 0;
  void f(const double* t1, double* t2) {   // em.rs:75-117
    if (nclasses) k_class_norm<<<(unsigned)((nclasses * 32 + 255) / 256), 256, 0, st>>>(off.p, tx.p, cnt.p, nclasses, t1, norm.p, nullptr);
    k_tx_gather<<<(unsigned)(((uint64_t)nitems * 32 + 255) / 256), 256, 0, st>>>(toff.p, tcls.p, cnt.p, nitems, t1, norm.p, t2);
    k_sum2<<<1, 1024, 0, st>>>(t2, nullptr, nitems, scal.p);
    k_scale_by_inv<<<(nitems + 255) / 256, 256, 0, st>>>(t2, nitems, scal.p, 0);
    launches += 4;
  }
  double likelihood(const double* th) {    // em.rs:119-137
    if (!nclasses) return 0.0;
    k_class_norm<<<(unsigned)((nclasses * 32 + 255) / 256), 256, 0, st>>>(off.p, tx.p, cnt.p, nclasses, th, nullptr, llt.p);
    k_sum2<<<1, 1024, 0, st>>>(llt.p, nullptr, nclasses, scal.p + 2);
    launches += 2;
    double ll[2];
    HLAC_CUDA(cudaMemcpyAsync(ll, scal.p + 2, sizeof ll, cudaMemcpyDeviceToHost, st));
    HLAC_CUDA(cudaStreamSynchronize(st));
    return ll[0];
  }
};

}  // namespace

extern "C" {

int hlac_squarem(const hlac_class_tables* t, int device, double* theta_out, uint32_t* iters_out) try {
  if (!t || !theta_out) throw Error(HLAC_ERR_ARG, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw Error(HLAC_ERR_CUDA, "no CUDA device available (no CPU fallback)");
  if (device < 0 || device >= ndev) throw Error(HLAC_ERR_ARG, "bad device ordinal");
  DeviceGuard g(device);
  const uint32_t n = t->nitems;
  if (n == 0) { if (iters_out) *iters_out = 0; return HLAC_OK; }
  const uint64_t nc = t->n_umi_classes;
  const uint64_t nnz = nc ? t->umi_off[nc] : 0;
  cudaStream_t st; HLAC_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } sg{st};
  EmDevice em{}; em.st = st; em.nitems = n; em.nclasses = nc; em.nnz = nnz;
  // transposed CSR: tx -> classes containing it (class order ascending => fixed summation order)
  std::vector<uint64_t> toff(n + 1, 0); std::vector<uint32_t> tcls(std::max<uint64_t>(nnz, 1));
  for (uint64_t q = 0; q < nnz; q++) { if (t->umi_tx[q] >= n) throw Error(HLAC_ERR_ARG, "tx id out of range"); toff[t->umi_tx[q] + 1]++; }
  for (uint32_t i = 0; i < n; i++) toff[i + 1] += toff[i];
  { std::vector<uint64_t> cur(toff.begin(), toff.end() - 1);
    for (uint64_t c = 0; c < nc; c++) for (uint64_t q = t->umi_off[c]; q < t->umi_off[c + 1]; q++) tcls[cur[t->umi_tx[q]]++] = (uint32_t)c; }
  static const uint64_t zero_off[1] = {0}; static const uint32_t zero32[1] = {0};
  em.off.upload(nc ? t->umi_off : zero_off, nc + 1, st);
  em.tx.upload(nnz ? t->umi_tx : zero32, std::max<uint64_t>(nnz, 1), st);
  em.cnt.upload(nc ? t->umi_cnt : zero32, std::max<uint64_t>(nc, 1), st);
  em.toff.upload(toff.data(), n + 1, st); em.tcls.upload(tcls.data(), tcls.size(), st);
  em.norm.reserve(std::max<uint64_t>(nc, 1), st); em.llt.reserve(std::max<uint64_t>(nc, 1), st); em.raw.reserve(n, st); em.scal.reserve(8, st);
  DevBuf<double> th0, th1, th2, thsq, r, v, rsq, vsq;
  for (auto* b : {&th0, &th1, &th2, &thsq, &r, &v, &rsq, &vsq}) b->reserve(n, st);
  std::vector<double> init(n, 1.0 / (double)n);   // em.rs:49-51
  th0.upload(init.data(), n, st);
  const unsigned gb = (n + 255) / 256;
  uint32_t iters = 0;
  double* p0 = th0.p; double* p1 = th1.p; double* p2 = th2.p; double* psq = thsq.p;
  for (;;) {
    em.f(p0, p1); em.f(p1, p2);
    k_rv<<<gb, 256, 0, st>>>(p0, p1, p2, n, r.p, v.p, rsq.p, vsq.p);
    k_sum2<<<1, 1024, 0, st>>>(rsq.p, vsq.p, n, em.scal.p + 4);
    em.launches += 2;
    double sq[2];
    HLAC_CUDA(cudaMemcpyAsync(sq, em.scal.p + 4, sizeof sq, cudaMemcpyDeviceToHost, st));
    HLAC_CUDA(cudaStreamSynchronize(st));
    double alpha = -std::sqrt(sq[0]) / std::sqrt(sq[1]);   // em.rs:263 (inf/NaN when vsq == 0, as in the reference)
    int tries = 1;
    double lsq, l2;
    l2 = em.likelihood(p2);
    for (;;) {
      k_theta_sq<<<gb, 256, 0, st>>>(p0, r.p, v.p, n, alpha, psq);
      k_sum2<<<1, 1024, 0, st>>>(psq, nullptr, n, em.scal.p + 6);
      k_scale_by_inv<<<gb, 256, 0, st>>>(psq, n, em.scal.p + 6, 1);
      em.launches += 3;
      lsq = em.likelihood(psq);
      if (lsq > l2 || tries > 5) break;
      alpha = (alpha + -1.0) / 2.0;
      tries++;
    }
    double* chosen = lsq > l2 ? psq : p2;
    k_diffs<<<1, 1024, 0, st>>>(p0, chosen, n, em.scal.p);
    em.launches++;
    double d[2];
    HLAC_CUDA(cudaMemcpyAsync(d, em.scal.p, sizeof d, cudaMemcpyDeviceToHost, st));
    HLAC_CUDA(cudaStreamSynchronize(st));
    if (lsq > l2) std::swap(p0, psq); else std::swap(p0, p2);
    iters++;
    if ((d[1] < EM_ABS_TH && d[0] < EM_REL_TH) || iters > EM_ITERS) break;
  }
  HLAC_CUDA(cudaMemcpyAsync(theta_out, p0, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  HLAC_CUDA(cudaStreamSynchronize(st));
  if (iters_out) *iters_out = iters;
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

// The caller (em.rs:361-434) on explicit tables: pair_reads_explained from the counts_reads CSR.
int hlac_call_genotypes(const hlac_index* idx, const hlac_class_tables* t, const double* theta, hlac_calls* out) try {
  if (!idx || !t || !theta || !out) throw Error(HLAC_ERR_ARG, "null argument");
  const auto& names = idx->host.tx_names;
  if (names.size() != t->nitems) throw Error(HLAC_ERR_ARG, "class tables do not match the index");
  if (t->n_read_classes && !t->reads_off) throw Error(HLAC_ERR_ARG, "tables were produced in lean mode (no counts_reads): use hlac_geno_call on the session");
  const Ranked rk = rank_alleles(names, t->nitems, t->reads_explained, theta);
  // pair_reads_explained (em.rs:36-44) = reads in classes containing a or b. The reference rescans every class per
  // pair; here one pass builds, for the few candidate alleles only (<= WEIGHTS_TO_OUTPUT per gene), the list of
  // classes containing each, and a pair is the read total over the union of the two ascending class lists.
  std::vector<std::vector<uint32_t>> cls_of(rk.cands.size());
  for (uint64_t c = 0; c < t->n_read_classes; c++)
    for (uint64_t q = t->reads_off[c]; q < t->reads_off[c + 1]; q++) {
      const int32_t k = rk.cand_of[t->reads_tx[q]];
      if (k >= 0 && (cls_of[k].empty() || cls_of[k].back() != (uint32_t)c)) cls_of[k].push_back((uint32_t)c);
    }
  auto pair_explained = [&](uint32_t a, uint32_t b) {
    const auto& la = cls_of[rk.cand_of[a]]; const auto& lb = cls_of[rk.cand_of[b]];
    uint64_t e = 0; size_t i = 0, j = 0;
    while (i < la.size() || j < lb.size()) {   // union of two ascending class lists
      uint32_t c;
      if (j >= lb.size() || (i < la.size() && la[i] < lb[j])) c = la[i++];
      else if (i >= la.size() || lb[j] < la[i]) c = lb[j++];
      else { c = la[i]; i++; j++; }
      e += t->reads_cnt[c];
    }
    return (uint32_t)e;
  };
  call_from_ranked(rk, t->nreads, pair_explained, out);
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_calls_free(hlac_calls* c) {
  if (c) { free(c->called); free(c->w_tx); free(c->w_em); free(c->w_frac); free(c->p_a); free(c->p_b); free(c->p_frac); memset(c, 0, sizeof *c); }
  return HLAC_OK;
}

}  // extern "C"
