// Device-side building blocks of the batched ADMM solve (internal header): OSQP constants, scalar helpers, the TMA record copy,
// the cross-stage communication layer, small dense blocks, the per-stage register state and the scratch-line row map.
// The kernel itself is in admm_kernel_impl.cuh.
#pragma once
#include "admm_kernel.cuh"

namespace f110 {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr double OSQP_INFTY = 1e30;
constexpr double RHO_MIN = 1e-6, RHO_MAX = 1e6, RHO_EQ_OVER_RHO_INEQ = 1e3, RHO_TOL = 1e-4;
constexpr double MIN_SCALING = 1e-4, MAX_SCALING = 1e4;
constexpr double INF_THRESH = OSQP_INFTY * MIN_SCALING;  // 1e26

enum : int {
  ST_SOLVED = 1, ST_SOLVED_INACC = 2, ST_PINF_INACC = 3, ST_DINF_INACC = 4,
  ST_MAX_ITER = -2, ST_PINF = -3, ST_DINF = -4, ST_NON_CVX = -7, ST_UNSOLVED = -10
};

// plain compare-select min/max: every operand here is finite, so fmax/fmin's NaN handling is dead weight
__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }
// Reductions over the G-lane group of the calling lane; `mask` names exactly that group's lanes, so groups of one warp may
// sit in different branches (a QP that needs an infeasibility test next to one that does not).
// Max over the group.  Every caller reduces norms (non-negative values): for those the IEEE-754 order is the order of the bit
// patterns as unsigned integers, so the hardware integer warp reduction (REDUX) does it in two steps — the high words, then the
// low words of the lanes that hold the winning high word — instead of a five-step shuffle butterfly.  A NaN pattern sorts above
// every finite value and therefore propagates.
// (Lane groups narrower than the warp keep the butterfly: REDUX would run once per distinct member mask.)
template <int G = 32>
__device__ __forceinline__ double wmax(double v, unsigned mask = FULL) {
  if constexpr (G == 32) {
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned H = __reduce_max_sync(FULL, hi);
    const unsigned L = __reduce_max_sync(FULL, hi == H ? lo : 0u);
    return __hiloint2double((int)H, (int)L);
  } else {
#pragma unroll
    for (int o = G / 2; o; o >>= 1) v = dmax(v, __shfl_xor_sync(mask, v, o));
    return v;
  }
}
template <int G = 32>
__device__ __forceinline__ double wsum(double v, unsigned mask = FULL) {
#pragma unroll
  for (int o = G / 2; o; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}
__device__ __forceinline__ double limit_scaling(double v) {
  v = v < MIN_SCALING ? 1.0 : v;
  return v > MAX_SCALING ? MAX_SCALING : v;
}
// 1/sqrt(x) for x in [1e-4, 1e4] (the range limit_scaling leaves): MUFU seed y on the high word (about 2^-21 relative, no float
// round trip), then ONE third-order step  y (1 + E/2 + 3 E^2/8),  E = 1 - x y^2  (next term 5 E^3/16 < 1e-18): five FP64
// instructions instead of the seven of two Newton steps.  Within ~2 ulp of 1.0 / sqrt(x), which is all the Ruiz vectors need.
__device__ __forceinline__ double rsqrt_scaling(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double E = fma(-(x * y), y, 1.0);
  return fma(y * E, fma(0.375, E, 0.5), y);
}
__device__ __forceinline__ double clampd(double v, double lo, double hi) { return dmin(dmax(v, lo), hi); }
// 1/x for the positive, well-scaled pivots and rho values of the factor step: MUFU seed y (about 2^-20 relative) + one third-order
// step  y (1 + e + e^2),  e = 1 - x y  (next term e^3 < 1e-18) -> within an ulp of the quotient, without the special-case tail of a
// full division.
__device__ __forceinline__ double rcp_pos(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  return fma(y, fma(e, e, e), y);
}

// ---- bulk asynchronous copy (TMA) of the parameter record into shared memory ------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(arrivals));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar),
      "r"(phase)
      : "memory");
}

// ---- tensor memory (TMEM) as a per-lane store for the PCR multipliers ---------------------------------------------
// The multipliers are written once per factor step and read once per ADMM iteration, by the lane that wrote them, and nothing
// else touches them: shared memory is the wrong place for that — at 128 B/clk/SM it was the round-1 kernel's binding resource
// (44 LDS.128 = 176 wavefronts per warp-iteration).  Tensor memory is 128 lanes x 512 32-bit columns per SM with its own read
// path (tcgen05.ld, measured 64 B/clk per SM sub-partition, scripts/ubench/tmem_bw.cu); shape .32x32b gives lane i of warp w
// the row 32 (w % 4) + i, so one warp of a four-warp CTA owns a quarter of the CTA's columns x 32 rows: a private strip of
// `cols` x 4 bytes per lane.  No tensor-core instruction is involved.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t cols) {   // one full warp; cols a power of two >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t addr, uint32_t cols) {          // the warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_pair(uint32_t addr, double a, double b) {   // one 16-byte pair -> 4 columns
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(__double2loint(a)), "r"(__double2hiint(a)),
               "r"(__double2loint(b)), "r"(__double2hiint(b))
               : "memory");
}
// Asynchronous loads of NP consecutive pairs (NP = 1, 2, 4 or 8 -> .x4 / .x8 / .x16 / .x32).  The registers are valid only after
// tmem_wait_ld(); tmem_tie() then makes every consumer depend on that wait (the compiler knows nothing about the asynchrony).
__device__ __forceinline__ void tmem_ld1(uint32_t addr, double2* o) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
  o[0] = make_double2(__hiloint2double((int)r[1], (int)r[0]), __hiloint2double((int)r[3], (int)r[2]));
}
__device__ __forceinline__ void tmem_ld2(uint32_t addr, double2* o) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(addr));
#pragma unroll
  for (int j = 0; j < 2; ++j) o[j] = make_double2(__hiloint2double((int)r[4 * j + 1], (int)r[4 * j]), __hiloint2double((int)r[4 * j + 3], (int)r[4 * j + 2]));
}
__device__ __forceinline__ void tmem_ld4(uint32_t addr, double2* o) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(addr));
#pragma unroll
  for (int j = 0; j < 4; ++j) o[j] = make_double2(__hiloint2double((int)r[4 * j + 1], (int)r[4 * j]), __hiloint2double((int)r[4 * j + 3], (int)r[4 * j + 2]));
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, double2* o) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(addr));
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = make_double2(__hiloint2double((int)r[4 * j + 1], (int)r[4 * j]), __hiloint2double((int)r[4 * j + 3], (int)r[4 * j + 2]));
}
template <int NP>
__device__ __forceinline__ void tmem_ld_pairs(uint32_t addr, double2* o) {   // pair j sits in columns 4 j .. 4 j + 3
  if constexpr (NP >= 8) { tmem_ld8(addr, o); tmem_ld_pairs<NP - 8>(addr + 32, o + 8); }
  else if constexpr (NP >= 4) { tmem_ld4(addr, o); tmem_ld_pairs<NP - 4>(addr + 16, o + 4); }
  else if constexpr (NP >= 2) { tmem_ld2(addr, o); tmem_ld_pairs<NP - 2>(addr + 8, o + 2); }
  else if constexpr (NP == 1) { tmem_ld1(addr, o); }
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
template <int NP>
__device__ __forceinline__ void tmem_tie(double2* v) {
#pragma unroll
  for (int j = 0; j < NP; ++j) asm volatile("" : "+d"(v[j].x), "+d"(v[j].y));
}

// ---- cross-stage communication ------------------------------------------------------------------------------
// One stage per thread.  WPQ = warps per QP: 1 -> everything is a warp shuffle; 2 or 4 (horizons 32..127) -> the
// CTA is the QP, values travel through a double-buffered shared-memory exchange with one barrier per exchange
// (a thread can only overwrite buffer b after passing the barrier of the exchange on buffer b^1, which every
// thread reaches only after it has finished reading b).
// G = lanes per QP when several short-horizon QPs share one warp (WPQ == 1 only): shuffles are confined to the
// G-lane segment (the width argument), and every shuffle / vote names only the group's lanes in its mask.
template <int WPQ, int G = 32>
struct Comm;

template <int G>
struct Comm<1, G> {
  unsigned gmask;   // lanes of this lane's group
  __device__ __forceinline__ Comm(double*, int tid, int = 0)
      : gmask(G == 32 ? FULL : (((1u << (G & 31)) - 1u) << ((unsigned)tid & ~(unsigned)(G - 1) & 31u))) {}
  __device__ __forceinline__ unsigned m() const { return G == 32 ? FULL : gmask; }
  template <int K> __device__ __forceinline__ void up(const double* v, double* o, int h) {
#pragma unroll
    for (int i = 0; i < K; ++i) o[i] = __shfl_up_sync(m(), v[i], h, G);
  }
  template <int K> __device__ __forceinline__ void dn(const double* v, double* o, int h) {
#pragma unroll
    for (int i = 0; i < K; ++i) o[i] = __shfl_down_sync(m(), v[i], h, G);
  }
  template <int K> __device__ __forceinline__ void both(const double* v, double* lo, double* hi, int h) {
#pragma unroll
    for (int i = 0; i < K; ++i) { lo[i] = __shfl_up_sync(m(), v[i], h, G); hi[i] = __shfl_down_sync(m(), v[i], h, G); }
  }
  template <int K> __device__ __forceinline__ void xr(const double* v, double* o, int h) {  // partner k ^ h
#pragma unroll
    for (int i = 0; i < K; ++i) o[i] = __shfl_xor_sync(m(), v[i], h, G);
  }
  __device__ __forceinline__ double rmax(double v) { return wmax<G>(v, m()); }
  __device__ __forceinline__ double rsum(double v) { return wsum<G>(v, m()); }
  __device__ __forceinline__ bool any(bool b) { return __any_sync(m(), b); }
  // KM maxima and KS sums over the group, in place (one exchange when the QP spans several warps)
  template <int KM, int KS> __device__ __forceinline__ void reduce(double* mx, double* sm) {
#pragma unroll
    for (int i = 0; i < KM; ++i) mx[i] = rmax(mx[i]);
#pragma unroll
    for (int j = 0; j < KS; ++j) sm[j] = rsum(sm[j]);
  }
  __device__ __forceinline__ void sync() { __syncwarp(); }
};

template <int WPQ, int G>
struct Comm {
  static constexpr int T = 32 * WPQ;
  static constexpr int KMAX = (WPQ == 4 || ADMM_W2_GLOBAL_LEVELS > 0) ? 5 : 9;   // values per exchange; with four warps the 9-wide ones (factor step only) go in two rounds to save shared memory
  double* xb;   // [2][KMAX][T] exchange buffers
  double* rb;   // [2][WPQ] reduction slots
  double* bb;   // [2][4][4] boundary slots (partitioned solve)
  double* kb;   // [2][RK][WPQ] slots of the several-values-at-once reduction
  static constexpr int RK = 16;
  int tid, xph = 0, rph = 0, bph = 0, kph = 0;
  int bar;      // hardware barrier of this QP's T threads: 0 when the QP is the whole CTA, 1 + q when several QPs share a CTA
  __device__ __forceinline__ Comm(double* smem, int t, int bar_id = 0)
      : xb(smem), rb(smem + 2 * KMAX * T), bb(smem + 2 * KMAX * T + 2 * WPQ), kb(smem + 2 * KMAX * T + 2 * WPQ + 32), tid(t), bar(bar_id) {}
  __device__ __forceinline__ void barrier() const {   // immediate barrier numbers: a register operand would make ptxas reserve all 16
    if (bar == 0) asm volatile("bar.sync 0, %0;" ::"n"(T) : "memory");
    else if (bar == 1) asm volatile("bar.sync 1, %0;" ::"n"(T) : "memory");
    else asm volatile("bar.sync 2, %0;" ::"n"(T) : "memory");
  }
  static constexpr int doubles() { return 2 * KMAX * T + 2 * WPQ + 32 + 2 * RK * WPQ; }
  // (Measured and dropped for two-warp QPs: shifts by one as a warp shuffle + one boundary value through a slot guarded by a one-way
  //  hardware barrier — bar.arrive by the producing warp, bar.sync by the consuming one.  Fewer stalls, but 60 more instructions per
  //  iteration: N=50 0.575 vs 0.521 ms per 4096 QPs.)
  // KM maxima and KS sums over the QP, in place, with ONE barrier: warp-level reduction, one slot per value and warp, then every
  // thread combines the warps' partial results in warp order (the same arithmetic as rmax / rsum one value at a time).
  template <int KM, int KS> __device__ __forceinline__ void reduce(double* mx, double* sm) {
    static_assert(KM + KS <= RK, "more values than reduction slots");
    double* r = kb + kph * RK * WPQ;
    kph ^= 1;
#pragma unroll
    for (int i = 0; i < KM; ++i) mx[i] = wmax<32>(mx[i]);
#pragma unroll
    for (int j = 0; j < KS; ++j) sm[j] = wsum<32>(sm[j]);
    if ((tid & 31) == 0) {
#pragma unroll
      for (int i = 0; i < KM; ++i) r[i * WPQ + (tid >> 5)] = mx[i];
#pragma unroll
      for (int j = 0; j < KS; ++j) r[(KM + j) * WPQ + (tid >> 5)] = sm[j];
    }
    barrier();
#pragma unroll
    for (int i = 0; i < KM; ++i) {
      double m = r[i * WPQ];
#pragma unroll
      for (int w = 1; w < WPQ; ++w) m = dmax(m, r[i * WPQ + w]);
      mx[i] = m;
    }
#pragma unroll
    for (int j = 0; j < KS; ++j) {
      double m = r[(KM + j) * WPQ];
#pragma unroll
      for (int w = 1; w < WPQ; ++w) m += r[(KM + j) * WPQ + w];
      sm[j] = m;
    }
  }
  // Partitioned solve: the two stages on either side of a partition boundary post a 3-vector; every thread gets the value of its own
  // side (`own`) and of the other side (`other`).  LEVEL 1: warps (0, 1) — and (2, 3) of a four-warp QP — across stages 31 | 32
  // (95 | 96); LEVEL 2 (four warps): the two pairs across stages 63 | 64.  Double-buffered like the exchanges above.
  template <int LEVEL> __device__ __forceinline__ void boundary(const double* v, double* own, double* other) {
    static_assert((WPQ == 2 && LEVEL == 1) || (WPQ == 4 && (LEVEL == 1 || LEVEL == 2)), "boundary slots: two or four warps per QP");
    double* b = bb + bph * 16;
    bph ^= 1;
    const int w = tid >> 5;
    // slot of this thread's side: LEVEL 1 -> one per warp; LEVEL 2 -> one per pair of warps
    const int side = LEVEL == 1 ? w : (w >> 1);
    const bool posts = LEVEL == 1 ? ((tid & 63) == 31 || (tid & 63) == 32) : (tid == 63 || tid == 64);
    if (posts) { b[4 * side] = v[0]; b[4 * side + 1] = v[1]; b[4 * side + 2] = v[2]; }
    barrier();
#pragma unroll
    for (int i = 0; i < 3; ++i) { own[i] = b[4 * side + i]; other[i] = b[4 * (side ^ 1) + i]; }
  }
  template <int K> __device__ __forceinline__ double* put(const double* v) {
    static_assert(K <= KMAX, "exchange wider than the buffer");
    double* b = xb + xph * KMAX * T;
    xph ^= 1;
#pragma unroll
    for (int i = 0; i < K; ++i) b[i * T + tid] = v[i];
    barrier();
    return b;
  }
  // out-of-range sources return the caller's own value, like a shuffle; callers mask them
  template <int K> __device__ __forceinline__ void up(const double* v, double* o, int h) {
    if constexpr (K > KMAX) { up<KMAX>(v, o, h); up<K - KMAX>(v + KMAX, o + KMAX, h); }
    else {
      const double* b = put<K>(v);
      const int src = tid - h >= 0 ? tid - h : tid;
#pragma unroll
      for (int i = 0; i < K; ++i) o[i] = b[i * T + src];
    }
  }
  template <int K> __device__ __forceinline__ void dn(const double* v, double* o, int h) {
    if constexpr (K > KMAX) { dn<KMAX>(v, o, h); dn<K - KMAX>(v + KMAX, o + KMAX, h); }
    else {
      const double* b = put<K>(v);
      const int src = tid + h < T ? tid + h : tid;
#pragma unroll
      for (int i = 0; i < K; ++i) o[i] = b[i * T + src];
    }
  }
  template <int K> __device__ __forceinline__ void both(const double* v, double* lo, double* hi, int h) {
    if constexpr (K > KMAX) { both<KMAX>(v, lo, hi, h); both<K - KMAX>(v + KMAX, lo + KMAX, hi + KMAX, h); }
    else {
      const double* b = put<K>(v);
      const int sl = tid - h >= 0 ? tid - h : tid, sh = tid + h < T ? tid + h : tid;
#pragma unroll
      for (int i = 0; i < K; ++i) { lo[i] = b[i * T + sl]; hi[i] = b[i * T + sh]; }
    }
  }
  template <int K> __device__ __forceinline__ void xr(const double* v, double* o, int h) {
    if constexpr (K > KMAX) { xr<KMAX>(v, o, h); xr<K - KMAX>(v + KMAX, o + KMAX, h); }
    else {
      const double* b = put<K>(v);
      const int src = tid ^ h;
#pragma unroll
      for (int i = 0; i < K; ++i) o[i] = b[i * T + src];
    }
  }
  __device__ __forceinline__ double* rslot(double v) {
    double* r = rb + rph * WPQ;
    rph ^= 1;
    if ((tid & 31) == 0) r[tid >> 5] = v;
    barrier();
    return r;
  }
  __device__ __forceinline__ double rmax(double v) {
    const double* r = rslot(wmax<32>(v));
    double m = r[0];
#pragma unroll
    for (int w = 1; w < WPQ; ++w) m = dmax(m, r[w]);
    return m;
  }
  __device__ __forceinline__ double rsum(double v) {
    const double* r = rslot(wsum<32>(v));
    double m = r[0];
#pragma unroll
    for (int w = 1; w < WPQ; ++w) m += r[w];
    return m;
  }
  __device__ __forceinline__ bool any(bool b) {
    const double* r = rslot(__any_sync(FULL, b) ? 1.0 : 0.0);
    bool a = false;
#pragma unroll
    for (int w = 0; w < WPQ; ++w) a |= r[w] != 0.0;
    return a;
  }
  __device__ __forceinline__ void sync() { barrier(); }
};

// ---- 3x3 helpers (row-major double[9]) ------------------------------------------------------------
__device__ __forceinline__ void mm3(const double* a, const double* b, double* c) {  // c = a b
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
// inverse of a symmetric positive definite 3x3 via LDL^T (reads the lower triangle)
__device__ __forceinline__ void inv_spd3(const double* a, double* inv) {
  const double d0 = a[0], i0 = rcp_pos(d0);
  const double l10 = a[3] * i0, l20 = a[6] * i0;
  const double d1 = a[4] - l10 * a[3], i1 = rcp_pos(d1);
  const double l21 = (a[7] - l20 * a[3]) * i1;
  const double d2 = a[8] - l20 * a[6] - l21 * (a[7] - l20 * a[3]), i2 = rcp_pos(d2);
  const double m10 = -l10, m20 = l10 * l21 - l20, m21 = -l21;  // L^-1 = [[1,0,0],[m10,1,0],[m20,m21,1]]
  inv[0] = i0 + m10 * m10 * i1 + m20 * m20 * i2;
  inv[1] = inv[3] = m10 * i1 + m20 * m21 * i2;
  inv[2] = inv[6] = m20 * i2;
  inv[4] = i1 + m21 * m21 * i2;
  inv[5] = inv[7] = m21 * i2;
  inv[8] = i2;
}

// ---- D x D helpers for the steering-rate variant (row-major double[D*D]) ------------------------------
template <int D>
__device__ __forceinline__ void mmD(const double* a, const double* b, double* c) {  // c = a b
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      double v = a[D * i] * b[j];
#pragma unroll
      for (int t = 1; t < D; ++t) v = fma(a[D * i + t], b[D * t + j], v);
      c[D * i + j] = v;
    }
}
// inverse of a symmetric positive definite D x D via LDL^T (reads the lower triangle)
template <int D>
__device__ __forceinline__ void inv_spdD(const double* a, double* inv) {
  double L[D * D], dd[D], id[D], M[D * D];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    double dj = a[D * j + j];
#pragma unroll
    for (int t = 0; t < j; ++t) dj -= L[D * j + t] * L[D * j + t] * dd[t];
    dd[j] = dj;
    id[j] = rcp_pos(dj);
#pragma unroll
    for (int i = j + 1; i < D; ++i) {
      double v = a[D * i + j];
#pragma unroll
      for (int t = 0; t < j; ++t) v -= L[D * i + t] * L[D * j + t] * dd[t];
      L[D * i + j] = v * id[j];
    }
  }
  // M = L^-1 (unit lower triangular)
#pragma unroll
  for (int j = 0; j < D; ++j) {
    M[D * j + j] = 1.0;
#pragma unroll
    for (int i = j + 1; i < D; ++i) {
      double v = 0.0;
#pragma unroll
      for (int t = j; t < i; ++t) v -= L[D * i + t] * M[D * t + j];
      M[D * i + j] = v;
    }
  }
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i; j < D; ++j) {
      double v = 0.0;
#pragma unroll
      for (int t = j; t < D; ++t) v += M[D * t + i] * M[D * t + j] * id[t];
      inv[D * i + j] = v;
      inv[D * j + i] = v;
    }
}

struct Model {  // Model::Linearize output (model.cpp:30-59): A = I + [0 0 a02; 0 0 a12; 0 0 0], B = [b00 0; b10 0; b20 b21]
  double a02, a12, b00, b10, b20, b21;
};
__device__ __forceinline__ void A_mul(const Model& m, const double* v, double* o) {   // o = A v
  o[0] = v[0] + m.a02 * v[2]; o[1] = v[1] + m.a12 * v[2]; o[2] = v[2];
}
__device__ __forceinline__ void At_mul(const Model& m, const double* v, double* o) {  // o = A' v
  o[0] = v[0]; o[1] = v[1]; o[2] = m.a02 * v[0] + m.a12 * v[1] + v[2];
}
__device__ __forceinline__ void B_mul(const Model& m, const double* h, double* o) {   // o = B h
  o[0] = m.b00 * h[0]; o[1] = m.b10 * h[0]; o[2] = m.b20 * h[0] + m.b21 * h[1];
}
__device__ __forceinline__ void Bt_mul(const Model& m, const double* v, double* o) {  // o = B' v
  o[0] = m.b00 * v[0] + m.b10 * v[1] + m.b20 * v[2]; o[1] = m.b21 * v[2];
}

// Everything one lane keeps in registers for its stage.
struct Stage {
  // problem data
  double bd[3];            // dynamics rhs (l = u): -x_cur at k = 0, -C at k >= 1      (mpc.cpp:299,305)
  double gm[6];            // gap rows 2x3: ones at k = 0, [l1a l1b 0; l2a l2b 0] after (mpc.cpp:237-241, 260-272)
  double gl[2];            // gap lower bounds (upper is +INFTY)                       (mpc.cpp:279-300)
  double qx[3];            // -Q ref_k                                                 (mpc.cpp:225,228)
  // iterates (unscaled)
  double x[3], u[2];
  double zg[2], zb[2];      // (z of the dynamics rows is their right-hand side bd; the first iteration's value sits in the scratch line)
  double yd[3], yg[2], yb[2];
  // metric
  double sx[3], su[2];                 // sigma_j
  double rd[3], rg[2], rb[2];          // rho_i
  double rda[3];                       // alpha * rho_i of the dynamics rows
  double ig[2], ib[2];                 // 1 / rho_i (inequality rows only; equality rows project to l = u)
  // input elimination
  double wi[3];            // inverse of W_k = R + Sigma_u + rho_box + B' R_{k+1} B   (00, 01, 11)
  double rdn[3];           // rho of the NEXT stage's dynamics rows
  double rbm[4];           // R_{k+1} B: (rdn0 b00, rdn1 b10, rdn2 b20, rdn2 b21)
};

// What the steering-rate variant adds to a lane (empty otherwise).
template <bool RATE>
struct RateExt {};
template <>
struct RateExt<true> {
  double zr, yr;           // rate row iterate
  double rr, ir;           // its rho and 1/rho
  double rrn;              // rho of the NEXT stage's rate row
  double rbase;            // centre of its bounds: steer_prev on stage 0, else 0; the row lives in [rbase - D, rbase + D]
  double wvi;              // 1 / (R_v + sigma_v + rho_box_v + b_v' R_{k+1} b_v): pivot of the speed elimination
};
// What the state-box rows add to a lane (empty otherwise): three identity rows on x_k.
template <bool SBOX>
struct SBoxExt {};
template <>
struct SBoxExt<true> {
  double zs[3], ys[3];     // row iterates
  double rs[3], is[3];     // rho and 1/rho
  double slo[3], shi[3];   // bounds: x_cur -+ d on x and y, -+INFTY on the orientation (Constraints::SetXLims, constraints.cpp:108-114)
};
template <bool RATE, bool SBOX = false>
struct StageT : Stage, RateExt<RATE>, SBoxExt<SBOX> {};

// per-QP scratch line in global memory (L2): [SCR_ROWS_ALLOC][T] doubles, element-major, one column per stage
constexpr int SCR_DX = 0, SCR_DU = 3, SCR_ED = 5, SCR_EG = 8, SCR_EB = 10;       // scaling vectors D, E
constexpr int SCR_PX = 12, SCR_PU = 15, SCR_PYD = 17, SCR_PYG = 20, SCR_PYB = 22;  // iterate before the last step
constexpr int SCR_WD = 24, SCR_WG = 27, SCR_WB = 29;    // e_i^2 / c per row: rho_i = rho_bar_i * w_i (factor step only)
constexpr int SCR_CG = 31, SCR_CB = 33;                 // row class codes of the gap / box rows (factor step only)
constexpr int SCR_NQ = 35, SCR_SNQ = 36;                // ||q||_inf unscaled / scaled (termination checks only)
constexpr int SCR_ER = 37, SCR_WR = 38, SCR_CR = 39, SCR_PYR = 40;   // steering-rate row: E, e^2/c, class, previous y
constexpr int SCR_ZD = 41;                              // z of the dynamics rows before the first iteration
constexpr int SCR_ES = 44, SCR_WS = 47, SCR_CS = 50, SCR_PYS = 53;   // state-box rows: E, e^2/c, class, previous y
constexpr int SCR_ROWS = 56;
static_assert(SCR_ROWS <= SCR_ROWS_ALLOC, "scratch line too short");

}  // namespace

}  // namespace f110
