// SQUAREM EM over counts_umi on the GPU (src/em.rs:48-138 EmProblem for EqClassCounts, :232-311 squarem, :315-336
// diffs), f64 throughout, plus the genotype caller (src/em.rs:361-434) on the host.
//
// The EM step is a sparse gather/scatter: per class norm = sum(theta[tx]); per tx theta2[tx] += theta[tx]/norm*count.
// The reference iterates a std HashMap (summation order unspecified, hence the 1e-6 tolerance); here the scatter is a
// deterministic gather over the transposed CSR (one warp per tx, fixed tree order), so runs are reproducible.
// No tensor cores: nothing here is a dense contraction.

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "caller.hpp"
#include "device_index.hpp"
#include "radix_sort.cuh"
#include "scan.cuh"

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

// ---- the whole SQUAREM loop as ONE persistent cooperative kernel ---------------------------------------------------------
// The multi-kernel host loop below costs ~19 launches and 4 host round trips per outer iteration (0.9 ms at 30 k alleles,
// where one pass over the 12 M-entry class table is ~10 us of L2 traffic): launch- and sync-bound. Here one grid of
// co-resident blocks (cudaLaunchCooperativeKernel) runs every phase of every iteration and meets at a grid barrier
// between phases; scalars (totals, squared norms, likelihoods, diffs) are reduced deterministically - per-block partials
// in a fixed thread -> item mapping, then every block sums the partials in the same order - so that all blocks take the
// same branches of em.rs:263-307 without communicating the decision.
struct SqArgs {
  const uint64_t* off; const uint32_t* tx; const uint32_t* cnt;      // counts_umi CSR
  const uint64_t* toff; const uint32_t* tcls;                        // transposed: tx -> classes
  // both tables cut into segments of <= SQ_SEG entries so that one 30 k-member class (or one allele that is in 10 k
  // classes) is spread over many warps instead of serialising one: segment s covers [seg_begin[s], min(seg_begin[s] +
  // SQ_SEG, end of its row)); row r owns segments [row_seg[r], row_seg[r + 1])
  const uint64_t* cseg_begin; const uint32_t* cseg_row; const uint64_t* crow_seg; uint64_t n_cseg;
  const uint64_t* tseg_begin; const uint32_t* tseg_row; const uint64_t* trow_seg; uint64_t n_tseg;
  uint32_t n; uint64_t nc;
  double* th[4];                                                     // theta0, theta1, theta2, theta_sq (roles rotate)
  double* r; double* v; double* norm;
  double* cpart; double* tpart;                                      // per-segment partial sums
  double* part;                                                      // [2][gridDim.x] block partials
  unsigned* bar;                                                     // [0] arrivals, [1] generation
  uint32_t* out;                                                     // [0] iterations, [1] index of the final theta buffer
};
constexpr uint32_t SQ_SEG = 1024;

__device__ __forceinline__ void grid_barrier(unsigned* bar) {
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile unsigned* vgen = bar + 1;
    const unsigned gen = *vgen;
    __threadfence();
    if (atomicAdd(bar, 1u) == gridDim.x - 1) { bar[0] = 0; __threadfence(); atomicAdd(bar + 1, 1u); }
    else while (*vgen == gen) {}
    __threadfence();
  }
  __syncthreads();
}

// deterministic grid-wide sum (or max) of two per-thread values; contains one grid barrier; every thread of every block
// returns the same pair
template <bool MAX>
__device__ __forceinline__ void grid_reduce2(const SqArgs& a, double x, double y, double& ox, double& oy) {
  __shared__ double sx[32], sy[32], bx, by;
  auto comb = [](double p, double q) { if (MAX) return q > p ? q : p; return p + q; };
  for (int o = 16; o > 0; o >>= 1) { x = comb(x, __shfl_down_sync(0xFFFFFFFFu, x, o)); y = comb(y, __shfl_down_sync(0xFFFFFFFFu, y, o)); }
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) { sx[warp] = x; sy[warp] = y; }
  __syncthreads();
  if (warp == 0) {
    x = lane < nw ? sx[lane] : 0.0; y = lane < nw ? sy[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) { x = comb(x, __shfl_down_sync(0xFFFFFFFFu, x, o)); y = comb(y, __shfl_down_sync(0xFFFFFFFFu, y, o)); }
    if (lane == 0) { a.part[blockIdx.x] = x; a.part[gridDim.x + blockIdx.x] = y; }
  }
  grid_barrier(a.bar);
  if (warp == 0) {
    x = 0.0; y = 0.0;
    for (unsigned b = lane; b < gridDim.x; b += 32) { x = comb(x, __ldcg(a.part + b)); y = comb(y, __ldcg(a.part + gridDim.x + b)); }
    for (int o = 16; o > 0; o >>= 1) { x = comb(x, __shfl_down_sync(0xFFFFFFFFu, x, o)); y = comb(y, __shfl_down_sync(0xFFFFFFFFu, y, o)); }
    if (lane == 0) { bx = x; by = y; }
  }
  __syncthreads();
  ox = bx; oy = by;
  __syncthreads();   // bx/by and part[] are reused by the next reduction
}

// class norms norm[c] = sum theta[tx] (em.rs:79-82) in two balanced phases with a grid barrier between them: one warp
// per segment, then one thread per class adding its segments in order. Returns this thread's share of the
// log-likelihood sum count * ln(norm) over non-empty classes (em.rs:119-137) when asked. norm[] is complete for other
// blocks only after the caller's next grid barrier.
__device__ __forceinline__ double sq_class_norm(const SqArgs& a, const double* theta, bool want_ll) {
  const unsigned lane = threadIdx.x & 31;
  const uint64_t gt = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t sg = gt >> 5; sg < a.n_cseg; sg += nthreads >> 5) {
    const uint64_t b = __ldg(a.cseg_begin + sg), e = min(b + SQ_SEG, __ldg(a.off + __ldg(a.cseg_row + sg) + 1));
    double s = 0.0;
#pragma unroll 4
    for (uint64_t q = b + lane; q < e; q += 32) s += theta[__ldg(a.tx + q)];   // plain loads: L1 is invalidated by the fence in grid_barrier
    s = warp_sum(s);
    if (lane == 0) a.cpart[sg] = s;
  }
  grid_barrier(a.bar);
  double ll = 0.0;   // one warp per class, lanes over its segments: the partial loads are independent, not a latency chain
  for (uint64_t c = gt >> 5; c < a.nc; c += nthreads >> 5) {
    const uint64_t s0 = __ldg(a.crow_seg + c), s1 = __ldg(a.crow_seg + c + 1);
    double s = 0.0;
    for (uint64_t sg = s0 + lane; sg < s1; sg += 32) s += __ldcg(a.cpart + sg);
    s = warp_sum(s);
    if (lane == 0) { a.norm[c] = s; if (want_ll && s1 != s0) ll += (double)__ldg(a.cnt + c) * log(s); }
  }
  return ll;
}

// f() (em.rs:75-117) from theta `src` into `dst`: class norms | per-segment gather over the transposed table | per-allele
// totals and their sum | dst = raw / total. The caller places a barrier before dst is read by other blocks.
__device__ __forceinline__ void sq_f(const SqArgs& a, const double* src, double* dst, bool have_norm) {
  if (!have_norm) {   // norm[] already holds src's class norms when src was the theta_sq whose likelihood was just taken
    sq_class_norm(a, src, false);
    grid_barrier(a.bar);
  }
  const unsigned lane = threadIdx.x & 31;
  const uint64_t gt = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t sg = gt >> 5; sg < a.n_tseg; sg += nthreads >> 5) {
    const uint32_t t = __ldg(a.tseg_row + sg);
    const uint64_t b = __ldg(a.tseg_begin + sg), e = min(b + SQ_SEG, __ldg(a.toff + t + 1));
    const double th = __ldcg(src + t);
    double s = 0.0;
#pragma unroll 4
    for (uint64_t q = b + lane; q < e; q += 32) { const uint32_t c = __ldg(a.tcls + q); s += th / a.norm[c] * (double)__ldg(a.cnt + c); }
    s = warp_sum(s);
    if (lane == 0) a.tpart[sg] = s;
  }
  grid_barrier(a.bar);
  double mine = 0.0;
  for (uint64_t t = gt >> 5; t < a.n; t += nthreads >> 5) {
    const uint64_t s0 = __ldg(a.trow_seg + t), s1 = __ldg(a.trow_seg + t + 1);
    double s = 0.0;
    for (uint64_t sg = s0 + lane; sg < s1; sg += 32) s += __ldcg(a.tpart + sg);
    s = warp_sum(s);
    if (lane == 0) { dst[t] = s; mine += s; }
  }
  double total, unused;
  grid_reduce2<false>(a, mine, 0.0, total, unused);
  for (uint64_t t = gt >> 5; t < a.n; t += nthreads >> 5) if (lane == 0) dst[t] = dst[t] / total;   // the same lane wrote dst[t] above
}

__global__ void __launch_bounds__(1024, 1) k_squarem(const SqArgs a) {
  const uint64_t gt = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (uint64_t)gridDim.x * blockDim.x;
  int i0 = 0, i1 = 1, i2 = 2, isq = 3;
  uint32_t iters = 0;
  bool have_norm = false;   // uniform over the grid: every block takes the same branches
  for (;;) {
    double* t0 = a.th[i0]; double* t1 = a.th[i1]; double* t2 = a.th[i2]; double* tsq = a.th[isq];
    sq_f(a, t0, t1, have_norm);
    grid_barrier(a.bar);
    sq_f(a, t1, t2, false);
    grid_barrier(a.bar);
    double rs = 0.0, vs = 0.0;                                  // em.rs:255-261
    for (uint64_t i = gt; i < a.n; i += nthreads) {
      const double ri = __ldcg(t1 + i) - __ldcg(t0 + i), vi = __ldcg(t2 + i) - __ldcg(t1 + i) - ri;
      a.r[i] = ri; a.v[i] = vi; rs += ri * ri; vs += vi * vi;
    }
    double rsq, vsq;
    grid_reduce2<false>(a, rs, vs, rsq, vsq);
    double alpha = -sqrt(rsq) / sqrt(vsq);                      // em.rs:263
    double l2, lsq, unused;
    grid_reduce2<false>(a, sq_class_norm(a, t2, true), 0.0, l2, unused);
    int tries = 1;
    for (;;) {                                                  // em.rs:268-289
      const double asq = alpha * alpha;
      double part = 0.0;
      for (uint64_t i = gt; i < a.n; i += nthreads) {           // theta_sq = reg(t0 - 2 alpha r + alpha^2 v): clamp, then normalise
        double x = __ldcg(t0 + i) - 2.0 * alpha * a.r[i] + asq * a.v[i];
        if (x > 1.0) x = 1.0;
        if (x < 0.0) x = 1e-15;
        tsq[i] = x; part += x;
      }
      double nrm;
      grid_reduce2<false>(a, part, 0.0, nrm, unused);
      const double inv = 1.0 / nrm;
      for (uint64_t i = gt; i < a.n; i += nthreads) tsq[i] *= inv;   // own elements
      grid_barrier(a.bar);
      grid_reduce2<false>(a, sq_class_norm(a, tsq, true), 0.0, lsq, unused);
      if (lsq > l2 || tries > 5) break;
      alpha = (alpha + -1.0) / 2.0;
      tries++;
    }
    const double* chosen = lsq > l2 ? tsq : t2;
    double mr = 0.0, ma = 0.0;                                  // diffs() em.rs:315-336
    for (uint64_t i = gt; i < a.n; i += nthreads) {
      const double o = __ldcg(t0 + i), nw_ = __ldcg(chosen + i);
      const double d = fabs(o - nw_), rel = d / o;
      if (d > ma) ma = d;
      if (nw_ > EM_CARE_TH && rel > mr) mr = rel;
    }
    double rel_max, abs_max;
    grid_reduce2<true>(a, mr, ma, rel_max, abs_max);
    if (lsq > l2) { const int t = i0; i0 = isq; isq = t; } else { const int t = i0; i0 = i2; i2 = t; }
    have_norm = lsq > l2;   // the last class-norm pass ran on theta_sq, which is the next theta0
    iters++;
    if ((abs_max < EM_ABS_TH && rel_max < EM_REL_TH) || iters > EM_ITERS) break;
  }
  if (gt == 0) { a.out[0] = iters; a.out[1] = (uint32_t)i0; }
}

// transposed table on the device: entry q of class c becomes the pair (tx[q], c); a stable radix sort by tx leaves every
// allele's class list in ascending class order (= fixed summation order in the gather)
__global__ void k_entry_class(const uint64_t* __restrict__ off, uint64_t nc, uint32_t* cls_of_entry) {
  const uint64_t c = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= nc) return;
  for (uint64_t q = off[c] + (threadIdx.x & 31); q < off[c + 1]; q += 32) cls_of_entry[q] = (uint32_t)c;
}
__global__ void k_tx_hist(const uint32_t* __restrict__ tx, uint64_t nnz, uint32_t n, uint32_t* hist /* [n] */, unsigned int* bad) {
  const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nnz) return;
  const uint32_t t = tx[q];
  if (t >= n) { *bad = 1; return; }
  atomicAdd(hist + t, 1u);
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

static double em_now_s() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; }
#define EM_TICK(label) do { if (dbg) { HLAC_CUDA(cudaStreamSynchronize(st)); double t_ = em_now_s(); fprintf(stderr, "[hlac timing] em: %-24s %.3f s\n", label, t_ - t_last); t_last = t_; } } while (0)

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
  const bool dbg = getenv("HLAC_DEBUG_TIMING") != nullptr; double t_last = em_now_s();
  cudaStream_t st; HLAC_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } sg{st};
  EmDevice em{}; em.st = st; em.nitems = n; em.nclasses = nc; em.nnz = nnz;
  static const uint64_t zero_off[1] = {0}; static const uint32_t zero32[1] = {0};
  em.off.upload(nc ? t->umi_off : zero_off, nc + 1, st);
  em.tx.upload(nnz ? t->umi_tx : zero32, std::max<uint64_t>(nnz, 1), st);
  em.cnt.upload(nc ? t->umi_cnt : zero32, std::max<uint64_t>(nc, 1), st);
  EM_TICK("upload");
  // transposed CSR: tx -> classes containing it (class order ascending => fixed summation order), built on the device
  std::vector<uint64_t> toff(n + 1, 0);
  em.toff.reserve(n + 1, st); em.tcls.reserve(std::max<uint64_t>(nnz, 1), st);
  HLAC_CUDA(cudaMemsetAsync(em.toff.p, 0, ((size_t)n + 1) * 8, st));
  if (nnz) {
    DevBuf<uint32_t> cls_of_entry, cls_b, keys_a, keys_b, tx_hist; DevBuf<unsigned int> bad; DevBuf<uint64_t> scan_tmp; RadixSortScratch rs;
    cls_of_entry.reserve(nnz, st); cls_b.reserve(nnz, st); keys_a.reserve(nnz, st); keys_b.reserve(nnz, st); tx_hist.reserve(n, st); bad.reserve(1, st);
    HLAC_CUDA(cudaMemsetAsync(bad.p, 0, 4, st));
    HLAC_CUDA(cudaMemsetAsync(tx_hist.p, 0, (size_t)n * 4, st));
    k_entry_class<<<(unsigned)((nc * 32 + 255) / 256), 256, 0, st>>>(em.off.p, nc, cls_of_entry.p);
    k_tx_hist<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(em.tx.p, nnz, n, tx_hist.p, bad.p);
    HLAC_CUDA(cudaGetLastError());
    exclusive_scan_u32(tx_hist.p, em.toff.p, n, scan_tmp, st);   // toff[t] = entries of tx < t, toff[n] = nnz
    // (tx, class) pairs sorted by tx, stable: entries are generated in class order, so every tx's class list comes out ascending
    int bits = 1; while (bits < 32 && (1ull << bits) < n) bits++;
    HLAC_CUDA(cudaMemcpyAsync(keys_a.p, em.tx.p, nnz * 4, cudaMemcpyDeviceToDevice, st));
    const bool second = radix_sort_pairs(keys_a.p, keys_b.p, cls_of_entry.p, cls_b.p, (uint64_t)nnz, 0, bits, rs, st);
    HLAC_CUDA(cudaMemcpyAsync(em.tcls.p, second ? cls_b.p : cls_of_entry.p, nnz * 4, cudaMemcpyDeviceToDevice, st));
    unsigned int hbad = 0;
    HLAC_CUDA(cudaMemcpyAsync(&hbad, bad.p, 4, cudaMemcpyDeviceToHost, st));
    HLAC_CUDA(cudaMemcpyAsync(toff.data(), em.toff.p, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, st));
    HLAC_CUDA(cudaStreamSynchronize(st));
    if (hbad) throw Error(HLAC_ERR_ARG, "tx id out of range");
  }
  EM_TICK("transpose (device)");
  em.norm.reserve(std::max<uint64_t>(nc, 1), st); em.llt.reserve(std::max<uint64_t>(nc, 1), st); em.raw.reserve(n, st); em.scal.reserve(8, st);
  DevBuf<double> th0, th1, th2, thsq, r, v, rsq, vsq;
  for (auto* b : {&th0, &th1, &th2, &thsq, &r, &v, &rsq, &vsq}) b->reserve(n, st);
  std::vector<double> init(n, 1.0 / (double)n);   // em.rs:49-51
  th0.upload(init.data(), n, st);
  const char* mode = getenv("HLAC_EM");   // "host": the multi-kernel host loop below (kept for A/B runs and tests)
  if (!(mode && std::string(mode) == "host")) {
    int coop = 0, n_sm = 1, occ = 0;
    HLAC_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
    HLAC_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    HLAC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_squarem, 1024, 0));
    if (!coop || occ < 1) throw Error(HLAC_ERR_CUDA, "device cannot co-schedule the persistent SQUAREM kernel");
    // enough blocks to give every warp a few classes / alleles, never more than are co-resident
    const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)n_sm * occ, std::max<uint64_t>(1, (nnz / 32 + nc + n + 255) / 256));
    struct Segs { std::vector<uint64_t> begin, row_seg; std::vector<uint32_t> row; };
    auto cut = [](const uint64_t* off, uint64_t rows) {
      Segs sg; sg.row_seg.assign(rows + 1, 0);
      for (uint64_t r = 0; r < rows; r++) {
        for (uint64_t b = off[r]; b < off[r + 1]; b += SQ_SEG) { sg.begin.push_back(b); sg.row.push_back((uint32_t)r); }
        sg.row_seg[r + 1] = sg.begin.size();
      }
      return sg;
    };
    const Segs cs = cut(nc ? t->umi_off : zero_off, nc), ts = cut(toff.data(), n);
    DevBuf<uint64_t> d_cb, d_crs, d_tb, d_trs; DevBuf<uint32_t> d_cr, d_tr; DevBuf<double> cpart, tpart;
    static const uint64_t z64[1] = {0};
    d_cb.upload(cs.begin.empty() ? z64 : cs.begin.data(), std::max<size_t>(cs.begin.size(), 1), st);
    d_cr.upload(cs.row.empty() ? zero32 : cs.row.data(), std::max<size_t>(cs.row.size(), 1), st);
    d_crs.upload(cs.row_seg.data(), cs.row_seg.size(), st);
    d_tb.upload(ts.begin.empty() ? z64 : ts.begin.data(), std::max<size_t>(ts.begin.size(), 1), st);
    d_tr.upload(ts.row.empty() ? zero32 : ts.row.data(), std::max<size_t>(ts.row.size(), 1), st);
    d_trs.upload(ts.row_seg.data(), ts.row_seg.size(), st);
    cpart.reserve(std::max<size_t>(cs.begin.size(), 1), st); tpart.reserve(std::max<size_t>(ts.begin.size(), 1), st);
    DevBuf<double> part; DevBuf<unsigned> bar; DevBuf<uint32_t> outv;
    part.reserve(2 * (size_t)grid, st); bar.reserve(2, st); outv.reserve(2, st);
    HLAC_CUDA(cudaMemsetAsync(bar.p, 0, 2 * sizeof(unsigned), st));
    SqArgs a{};
    a.off = em.off.p; a.tx = em.tx.p; a.cnt = em.cnt.p; a.toff = em.toff.p; a.tcls = em.tcls.p; a.n = n; a.nc = nc;
    a.th[0] = th0.p; a.th[1] = th1.p; a.th[2] = th2.p; a.th[3] = thsq.p; a.r = r.p; a.v = v.p; a.norm = em.norm.p;
    a.cseg_begin = d_cb.p; a.cseg_row = d_cr.p; a.crow_seg = d_crs.p; a.n_cseg = cs.begin.size();
    a.tseg_begin = d_tb.p; a.tseg_row = d_tr.p; a.trow_seg = d_trs.p; a.n_tseg = ts.begin.size();
    a.cpart = cpart.p; a.tpart = tpart.p;
    a.part = part.p; a.bar = bar.p; a.out = outv.p;
    EM_TICK("segments");
    void* params[] = {&a};
    HLAC_CUDA(cudaLaunchCooperativeKernel((const void*)k_squarem, dim3(grid), dim3(1024), params, 0, st));
    uint32_t res[2] = {0, 0};
    HLAC_CUDA(cudaMemcpyAsync(res, outv.p, sizeof res, cudaMemcpyDeviceToHost, st));
    HLAC_CUDA(cudaStreamSynchronize(st));
    EM_TICK("k_squarem");
    if (res[1] > 3) throw Error(HLAC_ERR_STATE, "SQUAREM kernel returned a bad buffer index");
    HLAC_CUDA(cudaMemcpyAsync(theta_out, a.th[res[1]], (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    HLAC_CUDA(cudaStreamSynchronize(st));
    if (iters_out) *iters_out = res[0];
    return HLAC_OK;
  }
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
