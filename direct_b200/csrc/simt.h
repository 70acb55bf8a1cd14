// simt.h -- the thin layer the warp-per-trajectory solver is written against.
//
// The solver (ipddp_solver.h) is warp-synchronous code: 32 lanes own one trajectory, per-lane values
// live in registers (Reg<>), lanes talk through a per-warp shared-memory scratch separated by
// WARP_SYNC(), plus a handful of shuffle collectives.  Built by nvcc this maps 1:1 onto sm_100a SIMT.
// Built with -DDDP_EMULATE by a host compiler the same source runs the 32 lanes of each phase one
// after the other (FOR_LANES becomes a loop, Reg<> grows a lane dimension), which is how the kernel
// logic is debugged in a container without a GPU (tools/emulate.cpp, never shipped in the library).
#ifndef DIRECT_B200_SIMT_H_
#define DIRECT_B200_SIMT_H_

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__) && !defined(DDP_EMULATE)
#define DDP_GPU 1
#define DDP_DEVICE __device__ __forceinline__
#define DDP_DEVICE_NOINLINE __device__ __noinline__
#define DDP_HD __host__ __device__ inline
#define FOR_LANES(lane) for (int lane = lane_, once_ = 1; once_; once_ = 0)
#define WARP_SYNC() __syncwarp()
#define DDP_UNROLL _Pragma("unroll")
#define DDP_NOUNROLL _Pragma("unroll 1")
#define DDP_RESTRICT __restrict__
#else
#define DDP_GPU 0
#define DDP_DEVICE inline
#define DDP_DEVICE_NOINLINE inline
#define DDP_HD inline
#define FOR_LANES(lane) for (int lane = 0; lane < 32; ++lane)
#define WARP_SYNC() ((void)0)
#define DDP_UNROLL
#define DDP_NOUNROLL
#define DDP_RESTRICT __restrict__
#endif

namespace ddp {

// Per-lane register array: N values in every lane.
template <class T, int N> struct Reg {
#if DDP_GPU
    T v[N];
    DDP_DEVICE T &operator()(int, int i) { return v[i]; }
    DDP_DEVICE const T &operator()(int, int i) const { return v[i]; }
#else
    T v[32][N];
    T &operator()(int lane, int i) { return v[lane][i]; }
    const T &operator()(int lane, int i) const { return v[lane][i]; }
#endif
};

#if DDP_GPU
DDP_DEVICE double shfl_xor(double x, int m) { return __shfl_xor_sync(0xffffffffu, x, m); }
DDP_DEVICE float shfl_xor(float x, int m) { return __shfl_xor_sync(0xffffffffu, x, m); }
DDP_DEVICE double shfl_idx(double x, int s) { return __shfl_sync(0xffffffffu, x, s); }
DDP_DEVICE float shfl_idx(float x, int s) { return __shfl_sync(0xffffffffu, x, s); }
#endif

// A per-warp scratch pointer travels through structs and out-of-line calls as a generic pointer, and every access through it
// became a generic LD/ST plus address arithmetic (profiles/r1g: riccati had 352 LD.E.64 and not one LDS).  as_shared()
// re-derives the pointer from the dynamic shared-memory symbol, so the compiler knows the address space again.
#if DDP_GPU
template <class T> DDP_DEVICE T *as_shared(T *p) {
    extern __shared__ __align__(16) unsigned char ddp_dyn_smem[];
    const unsigned off = (unsigned)__cvta_generic_to_shared(p) - (unsigned)__cvta_generic_to_shared(ddp_dyn_smem);
    return reinterpret_cast<T *>(ddp_dyn_smem + off);
}
// The same for pointers into the workspace / the caller's arrays: tell the compiler they are global (LDG/STG).
template <class T> DDP_DEVICE T *as_global(T *p) {
    __builtin_assume(__isGlobal(p));
    return p;
}
#else
template <class T> DDP_DEVICE T *as_shared(T *p) { return p; }
template <class T> DDP_DEVICE T *as_global(T *p) { return p; }
#endif

// Butterfly all-reduce over the warp, element i of r; every lane ends with the same value, returned.
// The emulation reproduces the butterfly's association order so both builds round identically.
template <class T, int N> DDP_DEVICE T warp_sum(Reg<T, N> &r, int i, int lane_) {
#if DDP_GPU
    T x = r.v[i];
    DDP_UNROLL
    for (int m = 16; m >= 1; m >>= 1) x += shfl_xor(x, m);
    (void)lane_;
    return x;
#else
    (void)lane_;
    T t[32];
    for (int l = 0; l < 32; l++) t[l] = r.v[l][i];
    for (int m = 16; m >= 1; m >>= 1) {
        T u[32];
        for (int l = 0; l < 32; l++) u[l] = t[l] + t[l ^ m];
        for (int l = 0; l < 32; l++) t[l] = u[l];
    }
    return t[0];
#endif
}
template <class T, int N> DDP_DEVICE T warp_max(Reg<T, N> &r, int i, int lane_) {
#if DDP_GPU
    T x = r.v[i];
    DDP_UNROLL
    for (int m = 16; m >= 1; m >>= 1) { T y = shfl_xor(x, m); x = (y > x || y != y) ? y : x; }
    (void)lane_;
    return x;
#else
    (void)lane_;
    T x = r.v[0][i];
    int nan = 0;
    for (int l = 0; l < 32; l++) { T y = r.v[l][i]; if (y != y) nan = 1; if (y > x) x = y; }
    return nan ? (T)NAN : x;
#endif
}
// Value of element i held by lane `src` (warp-uniform result).
template <class T, int N> DDP_DEVICE T warp_bcast(const Reg<T, N> &r, int i, int src, int lane_) {
#if DDP_GPU
    (void)lane_;
    return shfl_idx(r.v[i], src);
#else
    (void)lane_;
    return r.v[src][i];
#endif
}
// Minimum over the warp of element i (int).
template <int N> DDP_DEVICE int warp_min_int(const Reg<int, N> &r, int i, int lane_) {
#if DDP_GPU
    (void)lane_;
    return __reduce_min_sync(0xffffffffu, r.v[i]);
#else
    (void)lane_;
    int x = r.v[0][i];
    for (int l = 1; l < 32; l++) if (r.v[l][i] < x) x = r.v[l][i];
    return x;
#endif
}
template <int N> DDP_DEVICE bool warp_any(const Reg<int, N> &r, int i, int lane_) {
#if DDP_GPU
    (void)lane_;
    return __any_sync(0xffffffffu, r.v[i]) != 0;
#else
    (void)lane_;
    for (int l = 0; l < 32; l++) if (r.v[l][i]) return true;
    return false;
#endif
}

// Asynchronous global -> shared copies (LDGSTS): the data never occupies registers, so a phase can keep several
// knots / row chunks in flight ahead of their use.  Per-thread semantics: cp_wait<N>() makes this thread's own
// copies (all but the N most recent groups) visible to itself; other lanes need a WARP_SYNC() after it.
#if DDP_GPU
template <int BYTES> DDP_DEVICE void cp_async(void *smem_dst, const void *gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gsrc), "n"(BYTES) : "memory");
}
DDP_DEVICE void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> DDP_DEVICE void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#else
template <int BYTES> DDP_DEVICE void cp_async(void *smem_dst, const void *gsrc) { memcpy(smem_dst, gsrc, BYTES); }
DDP_DEVICE void cp_commit() {}
template <int N> DDP_DEVICE void cp_wait() {}
#endif

// Bulk asynchronous copies (TMA, cp.async.bulk -> SASS UBLKCP) completing on an mbarrier: ONE lane issues the copy of a whole
// row / knot tile, the copy engine moves it while the warp computes, and the lanes wait on the barrier's phase parity right
// before they read the tile from shared memory.  Sizes are multiples of 16 bytes, both addresses 16-byte aligned.
// The CPU emulation copies at issue time; its waits are no-ops.
#if DDP_GPU
DDP_DEVICE void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
// make barrier initialisations / generic-proxy writes to shared memory visible to the async proxy (the copy engine)
DDP_DEVICE void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// ... and generic-proxy writes to global memory (slack rows written by st.global that a later bulk copy reads)
DDP_DEVICE void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
DDP_DEVICE void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
DDP_DEVICE void bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(__cvta_generic_to_global(gsrc)), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
DDP_DEVICE void mbar_wait(unsigned long long *bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(a), "r"(parity) : "memory");
}
#else
DDP_DEVICE void mbar_init(unsigned long long *, int) {}
DDP_DEVICE void fence_async_smem() {}
DDP_DEVICE void fence_async_all() {}
DDP_DEVICE void mbar_expect_tx(unsigned long long *, unsigned) {}
DDP_DEVICE void bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, unsigned long long *) { memcpy(smem_dst, gsrc, bytes); }
DDP_DEVICE void mbar_wait(unsigned long long *, unsigned) {}
#endif

// Two consecutive elements at a 2-element-aligned shared/global address as ONE access (128-bit in fp64, 64-bit in fp32).
template <class T> struct Pair2 { T x, y; };
#if DDP_GPU
DDP_DEVICE Pair2<double> ld2(const double *p) { const double2 v = *reinterpret_cast<const double2 *>(p); return {v.x, v.y}; }
DDP_DEVICE Pair2<float> ld2(const float *p) { const float2 v = *reinterpret_cast<const float2 *>(p); return {v.x, v.y}; }
DDP_DEVICE void st2(double *p, double a, double b) { *reinterpret_cast<double2 *>(p) = make_double2(a, b); }
DDP_DEVICE void st2(float *p, float a, float b) { *reinterpret_cast<float2 *>(p) = make_float2(a, b); }
#else
template <class T> DDP_DEVICE Pair2<T> ld2(const T *p) { return {p[0], p[1]}; }
template <class T> DDP_DEVICE void st2(T *p, T a, T b) { p[0] = a; p[1] = b; }
#endif

// log(): called rarely (LogProd) but ~100 SASS instructions per inlined fp64 copy; kept out of line so the hot
// row loops stay small (the v2 profile showed 35 % instruction-fetch stalls, profiles/r1b).
#if DDP_GPU
static __device__ __noinline__ double rlog(double x) { return log(x); }
#else
DDP_DEVICE double rlog(double x) { return log(x); }
#endif
DDP_DEVICE float rlog(float x) { return logf(x); }
// 1/x for the interior-point weights.  GPU fp64: MUFU.RCP64H seed + two Newton steps (~1 ulp, no slow-path
// call); arguments are slack/constraint values kept away from 0 by the fraction-to-boundary rule.
#if DDP_GPU
DDP_DEVICE double rrcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
DDP_DEVICE float rrcp(float x) { return 1.0f / x; }
#else
DDP_DEVICE double rrcp(double x) { return 1.0 / x; }
DDP_DEVICE float rrcp(float x) { return 1.0f / x; }
#endif
DDP_DEVICE double rsqrt_(double x) { return sqrt(x); }
DDP_DEVICE float rsqrt_(float x) { return sqrtf(x); }
// 1/sqrt(x): the Cholesky pivot scale.  GPU: rsqrt() (MUFU.RSQ64H + Newton, ~1 ulp), shorter dependent chain than
// sqrt followed by a division.
#if DDP_GPU
DDP_DEVICE double rrsqrt(double x) {   // x > 0 checked by the caller; seed ~2^-20, two Newton steps
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double h = 0.5 * x * y, e = fma(-h, y, 0.5);
    y = fma(y, e, y);
    h = 0.5 * x * y; e = fma(-h, y, 0.5);
    return fma(y, e, y);
}
DDP_DEVICE float rrsqrt(float x) { return rsqrtf(x); }
#else
DDP_DEVICE double rrsqrt(double x) { return 1.0 / sqrt(x); }
DDP_DEVICE float rrsqrt(float x) { return 1.0f / sqrtf(x); }
#endif
// unew = (uold + alpha ku) + Ku dx (ddp.cpp:689/:695).  The line search forms this value in the state recursion AND, to save
// registers, again in the row phase: one spelling (an explicit fused multiply-add on the GPU) so that both sites give the same bits.
#if DDP_GPU
DDP_DEVICE double step_u(double uo, double alpha, double k, double kdx) { return fma(alpha, k, uo) + kdx; }
DDP_DEVICE float step_u(float uo, float alpha, float k, float kdx) { return fmaf(alpha, k, uo) + kdx; }
#else
template <class T> DDP_DEVICE T step_u(T uo, T alpha, T k, T kdx) { return (uo + alpha * k) + kdx; }
#endif
DDP_DEVICE double rabs(double x) { return fabs(x); }
DDP_DEVICE float rabs(float x) { return fabsf(x); }
#if DDP_GPU
static __device__ __noinline__ double rpow(double x, double y) { return pow(x, y); }
#else
DDP_DEVICE double rpow(double x, double y) { return pow(x, y); }
#endif
DDP_DEVICE float rpow(float x, float y) { return powf(x, y); }
template <class T> DDP_DEVICE T rmax(T a, T b) { return (a < b) ? b : a; }   // std::max(a, b)
template <class T> DDP_DEVICE T rmin(T a, T b) { return (b < a) ? b : a; }   // std::min(a, b)
// max that propagates like Eigen's lpNorm<Infinity> over finite data; NaN in b is kept.
template <class T> DDP_DEVICE T amax(T a, T b) { return (b > a || b != b) ? b : a; }

}  // namespace ddp
#endif
