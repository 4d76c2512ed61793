// nda_device.cuh -- device-side building blocks shared by the pair kernels (sm_100a only).
//
//  * strict_*   : the reference's per-pair arithmetic with one IEEE rounding per operation
//                 (lib/openmp/src/getRadialNbrDensity.cpp:29-45,
//                 getWithin.cpp:58-77, getDistances.cpp:47-57).  Intrinsics (__fdiv_rn, __fmul_rn,
//                 ...) are used so nvcc can neither contract to FMA nor reassociate.
//  * fast_r2    : the FP32 filter applied to EVERY pair (two variants, see below).  It only decides
//                 which pairs are CANDIDATES (r2' < thr, thr = in-range limit plus a proven guard
//                 band, see prep_frames_kernel); candidates are re-evaluated with strict_* on the
//                 ORIGINAL coordinates, so every count / mask bit equals the reference's bit for bit.
//  * TMA        : 1-D bulk copies global -> shared (cp.async.bulk, SASS UBLKCP) completing on an
//                 mbarrier.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define NDA_FULL_MASK 0xffffffffu
#define NDA_MAGIC     12582912.0f          /* 1.5 * 2^23: x + MAGIC - MAGIC == rintf(x) for |x| < 2^22 */

struct __align__(16) FrameParams {
    float cx, cy, cz;        // cell edges of this frame
    float thr;               // candidate threshold on the fast r2 (valid only if exact_all == 0)
    float icx, icy, icz;     // fl(1/c)
    int   exact_all;         // 1: cut-off too close to half the box (or odd cell): every pair strict
};

// ---------------------------------------------------------------- strict arithmetic
__device__ __forceinline__ float strict_wrap(float d, float c)
{
    float q = __fdiv_rn(d, c);
    float n = roundf(q);                    // half away from zero, as C roundf
    float p = __fmul_rn(c, n);
    return __fsub_rn(d, p);
}

__device__ __forceinline__ float strict_norm2(float x, float y, float z)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

// a - b, wrapped, squared norm: bit-identical to the reference's float32 result
__device__ __forceinline__ float strict_r2(float ax, float ay, float az,
                                           float bx, float by, float bz,
                                           float cx, float cy, float cz)
{
    float wx = strict_wrap(__fsub_rn(ax, bx), cx);
    float wy = strict_wrap(__fsub_rn(ay, by), cy);
    float wz = strict_wrap(__fsub_rn(az, bz), cz);
    return strict_norm2(wx, wy, wz);
}

// (int)(sqrtf(r2)/dr) with the reference's range test 0 <= idx < outSize
// (getRadialNbrDensity.cpp:38-41).  NaN / huge values fail the test exactly as x86's
// cvttss2si result INT_MIN does.  Returns -1 when the pair is not counted.
__device__ __forceinline__ int strict_bin(float r2, float dr, int outSize)
{
    float t = __fdiv_rn(__fsqrt_rn(r2), dr);
    if (!(t > -1.0f && t < 2147483648.0f)) return -1;
    int idx = (int)t;                       // truncation toward zero
    return (idx >= 0 && idx < outSize) ? idx : -1;
}

// ---------------------------------------------------------------- strict results without the slow instructions
// The candidate path is hot in dense boxes (config 1: 46 % of the pairs are in range), so the three
// IEEE divisions d/c of the strict wrap get a division-free evaluation that is PROVABLY the same
// number, with the IEEE sequence as fallback whenever the proof's precondition fails.  (A bin-edge
// table replacing sqrt + division was tried as well and measured no gain: the drain is bound by
// the latency of its dependent chain, not by instruction count.)
//
//     n = roundf(fl(d/c)).  Let q' = fl(d * fl(1/c)), n' = rintf(q') (magic constant).  Both
//     q = fl(d/c) and q' are within 3.01 * 2^-24 relative of d/c, so |q - q'| <= 2^-22 |q'| <
//     2^-22 (|n'| + 1).  If |q' - n'| < 1/2 - 2^-21 (|n'| + 1) then q lies strictly inside
//     (n' - 1/2, n' + 1/2) and roundf(q) = n' (no tie can occur).  Otherwise (near a tie, NaN,
//     |q| >= 2^20, zero cell) the caller falls back to __fdiv_rn + roundf.
__device__ __forceinline__ bool wrap_checked(float d, float c, float ic, float &w)
{
    const float q = __fmul_rn(d, ic);
    const float n = __fadd_rn(__fadd_rn(q, NDA_MAGIC), -NDA_MAGIC);
    const float f = fabsf(__fsub_rn(q, n));                                  // exact
    const float bound = __fmaf_rn(fabsf(n), -4.76837158203125e-07f, 0.5f - 4.76837158203125e-07f);
    w = __fsub_rn(d, __fmul_rn(c, n));
    return f < bound;
}

// strict r2 of (a - b): the division-free route when all three axes pass the check, else IEEE
__device__ __forceinline__ float strict_r2_fastpath(float ax, float ay, float az, float bx, float by, float bz,
                                                    const FrameParams &fp)
{
    float wx, wy, wz;
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    const bool ok = wrap_checked(dx, fp.cx, fp.icx, wx) & wrap_checked(dy, fp.cy, fp.icy, wy) &
                    wrap_checked(dz, fp.cz, fp.icz, wz);
    if (!ok) {
        wx = strict_wrap(dx, fp.cx);
        wy = strict_wrap(dy, fp.cy);
        wz = strict_wrap(dz, fp.cz);
    }
    return strict_norm2(wx, wy, wz);
}

// ---------------------------------------------------------------- fast filters
// Two interchangeable FP32 filters, both applied to EVERY pair, both only deciding which pairs
// become candidates for the strict re-evaluation:
//   kFilterRint : original coordinates; n = rintf(d * (1/c)) via the magic constant, w = fma(-c, n, d)
//                 -> 15 FMA-pipe instructions per pair.
//   kFilterFold : coordinates pre-wrapped into [0, c) by pack_kernel, so d in (-c, c) and the
//                 minimum image is |w| = min(|d|, c - |d|): FADD, FADD, FMNMX per axis
//                 -> 9 FMA-pipe + 3 ALU-pipe instructions per pair (the two pipes run side by side).
//   kFilterFold2: the fold filter on two B atoms at a time with Blackwell's packed FP32 pipe
//                 instructions (sub/mul/fma.f32x2, SASS FADD2/FMUL2/FFMA2): the wrapped B tile is
//                 stored pair-interleaved [x0 x1 y0 y1 | z0 z1 w0 w1] so both operands of a packed
//                 instruction sit in adjacent registers -> ~9 issue slots per pair instead of 12.
constexpr int kFilterRint = 0;
constexpr int kFilterFold = 1;
constexpr int kFilterFold2 = 2;     // 2..5 = fold2 with 0..3 axes in the add form (see fold_axis2)
__host__ __device__ constexpr bool is_fold2(int f) { return f >= kFilterFold2; }

struct FrameRegs {
    float px, py, pz;        // rint: -c        fold: |c|
    float icx, icy, icz;     // rint: fl(1/c)   fold: |c|/2 (exact)
    float thr;
};

template <int kFilter>
__device__ __forceinline__ FrameRegs load_frame_regs(const FrameParams &p)
{
    FrameRegs r;
    if (kFilter == kFilterRint) { r.px = -p.cx; r.py = -p.cy; r.pz = -p.cz; }  // else: fold / fold2
    else                        { r.px = fabsf(p.cx); r.py = fabsf(p.cy); r.pz = fabsf(p.cz); }
    if (kFilter == kFilterRint) { r.icx = p.icx; r.icy = p.icy; r.icz = p.icz; }
    else { r.icx = 0.5f * fabsf(p.cx); r.icy = 0.5f * fabsf(p.cy); r.icz = 0.5f * fabsf(p.cz); }
    r.thr = p.thr;
    return r;
}

template <int kFilter>
__device__ __forceinline__ float fast_r2(float ax, float ay, float az, const float4 &b, const FrameRegs &f)
{
    const float dx = ax - b.x;
    const float dy = ay - b.y;
    const float dz = az - b.z;
    float wx, wy, wz;
    if (kFilter == kFilterRint) {
        const float nx = __fmaf_rn(dx, f.icx, NDA_MAGIC) - NDA_MAGIC;      // rintf(dx / cx)
        const float ny = __fmaf_rn(dy, f.icy, NDA_MAGIC) - NDA_MAGIC;
        const float nz = __fmaf_rn(dz, f.icz, NDA_MAGIC) - NDA_MAGIC;
        wx = __fmaf_rn(f.px, nx, dx);
        wy = __fmaf_rn(f.py, ny, dy);
        wz = __fmaf_rn(f.pz, nz, dz);
    } else {
        wx = fminf(fabsf(dx), f.px - fabsf(dx));
        wy = fminf(fabsf(dy), f.py - fabsf(dy));
        wz = fminf(fabsf(dz), f.pz - fabsf(dz));
    }
    return __fmaf_rn(wz, wz, __fmaf_rn(wy, wy, wx * wx));
}

// fold filter for one A atom against the B atoms (j, j+1) of a pair-interleaved tile:
// q0 = (x_j, x_j+1, y_j, y_j+1), q1 = (z_j, z_j+1, w_j, w_j+1).  Everything on the FMA pipe runs as
// packed f32x2 instructions (SASS FADD2 / FMUL2 / FFMA2; each half is rounded exactly like the
// scalar instruction, |x| and -x are free operand modifiers).  Per axis the folded component is
//   min form : e = c - |d| (FADD2), |w| = min(|d|, e)           (2 x FMNMX on the ALU pipe)
//   add form : t = |d| - c/2 (FADD2), |w| = c/2 - |t| (FADD2)     (no ALU instruction)
// kAddAxes of the three axes use the add form: it trades ALU-pipe cycles (half rate) for FMA-pipe
// cycles so that neither pipe saturates.
__device__ __forceinline__ void sub2_bcast_a(float a, float b0, float b1, float &d0, float &d1)
{
    asm("{\n\t.reg .b64 a, b, d;\n\t"
        "mov.b64 a, {%2, %2};\n\tmov.b64 b, {%3, %4};\n\t"
        "sub.f32x2 d, a, b;\n\tmov.b64 {%0, %1}, d;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a), "f"(b0), "f"(b1));
}

__device__ __forceinline__ void sub2_bcast_b(float a0, float a1, float b, float &d0, float &d1)
{
    asm("{\n\t.reg .b64 a, b, d;\n\t"
        "mov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %4};\n\t"
        "sub.f32x2 d, a, b;\n\tmov.b64 {%0, %1}, d;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b));
}

template <bool kAddForm>
__device__ __forceinline__ void fold_axis2(float a, float b0, float b1, float c, float h, float &w0, float &w1)
{
    float d0, d1;
    sub2_bcast_a(a, b0, b1, d0, d1);                       // d = a - b
    if (kAddForm) {
        float t0, t1;
        sub2_bcast_b(fabsf(d0), fabsf(d1), h, t0, t1);     // t = |d| - c/2
        sub2_bcast_a(h, fabsf(t0), fabsf(t1), w0, w1);     // |w| = c/2 - |t|
    } else {
        float e0, e1;
        sub2_bcast_a(c, fabsf(d0), fabsf(d1), e0, e1);     // e = c - |d|
        w0 = fminf(fabsf(d0), e0);
        w1 = fminf(fabsf(d1), e1);
    }
}

template <int kAddAxes>
__device__ __forceinline__ void fast_r2_fold2(float ax, float ay, float az, const float4 &q0, const float4 &q1,
                                              const FrameRegs &f, float &r0, float &r1)
{
    float wx0, wx1, wy0, wy1, wz0, wz1;
    fold_axis2<(kAddAxes >= 3)>(ax, q0.x, q0.y, f.px, f.icx, wx0, wx1);
    fold_axis2<(kAddAxes >= 2)>(ay, q0.z, q0.w, f.py, f.icy, wy0, wy1);
    fold_axis2<(kAddAxes >= 1)>(az, q1.x, q1.y, f.pz, f.icz, wz0, wz1);
    asm("{\n\t.reg .b64 x, y, z, r;\n\t"
        "mov.b64 x, {%2, %3};\n\tmov.b64 y, {%4, %5};\n\tmov.b64 z, {%6, %7};\n\t"
        "mul.f32x2 r, x, x;\n\tfma.rn.f32x2 r, y, y, r;\n\tfma.rn.f32x2 r, z, z, r;\n\t"
        "mov.b64 {%0, %1}, r;\n\t}"
        : "=f"(r0), "=f"(r1) : "f"(wx0), "f"(wx1), "f"(wy0), "f"(wy1), "f"(wz0), "f"(wz1));
}

// float index of component k (0..3 = x,y,z,w) of atom j in a pair-interleaved array
__device__ __forceinline__ size_t pair_layout_index(size_t j, int k)
{
    return (j >> 1) * 8 + (size_t)(2 * k) + (j & 1);
}

// ---------------------------------------------------------------- mbarrier + TMA bulk copy
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    uint32_t addr = smem_u32(bar);
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}

// global -> shared bulk copy (TMA, no tensor map needed for 1-D); bytes % 16 == 0, 16-B aligned
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
