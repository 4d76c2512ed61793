// nda_surface.cuh -- SURVEY.md section 8 row f-3: the two "water at the protein surface" loops of the
// reference that reuse the minimum-image inner loop with a different epilogue.
//
//   water_nearest_shift_kernel   lib/openmp/src/setWaterDistPBC.cpp:9-77      argmin epilogue
//   water_orient_kernel          lib/openmp/src/waterOrientAtSurface.cpp:10-112  top-maxN epilogue
//
// One thread owns one water (molecule) of one frame and walks the protein atoms IN INDEX ORDER (ties
// are resolved by order in the reference); protein tiles of 16-byte atoms are staged in shared
// memory by TMA bulk copies exactly like the B tiles of the pair kernels.  Both epilogues depend on
// the float32 distance of EVERY pair (running minimum / insertion order), so there is no filter
// here: every pair is evaluated with the strict arithmetic (division-free checked wrap with IEEE
// fallback, nda_device.cuh), which makes the results the reference's bit for bit by construction.
#pragma once
#include "nda_device.cuh"

constexpr int SURF_THREADS = 128;
constexpr int SURF_TILE = 1024;             // protein atoms per TMA stage
constexpr int SURF_MAXN = 16;               // upper bound on maxN (waterOrientAtSurface)

// w = d - c*roundf(d/c) AND the image vector p = c*roundf(d/c) (setWaterDistPBC.cpp:47-53)
__device__ __forceinline__ void strict_wrap_with_shift(float d, float c, float ic, float &w, float &p)
{
    const float q = __fmul_rn(d, ic);
    float n = __fadd_rn(__fadd_rn(q, NDA_MAGIC), -NDA_MAGIC);
    const float f = fabsf(__fsub_rn(q, n));
    const float bound = __fmaf_rn(fabsf(n), -4.76837158203125e-07f, 0.5f - 4.76837158203125e-07f);
    if (!(f < bound)) n = roundf(__fdiv_rn(d, c));            // near a tie / odd cell: the IEEE sequence
    p = __fmul_rn(c, n);
    w = __fsub_rn(d, p);
}

struct SurfTiles {
    float4 *sP;          // [2][SURF_TILE]
    uint64_t *mbar;      // [2]
};

// streams the protein tiles of one frame through shared memory and calls body(atomIndex, float4)
// for every protein atom in index order; all threads of the CTA must call it (barriers inside)
template <class Body>
__device__ __forceinline__ void for_each_protein_atom(const float4 *Pf, int nP, const SurfTiles &t,
                                                      unsigned &phase0, unsigned &phase1, Body body)
{
    const int tid = threadIdx.x;
    const int nTiles = (nP + SURF_TILE - 1) / SURF_TILE;
    auto tile_atoms = [&](int jt) { return min(SURF_TILE, nP - jt * SURF_TILE); };
    if (tid == 0) {
        const unsigned bytes = (unsigned)tile_atoms(0) * 16u;
        mbar_expect_tx(&t.mbar[0], bytes);
        tma_bulk_g2s(t.sP, Pf, bytes, &t.mbar[0]);
    }
    for (int jt = 0; jt < nTiles; ++jt) {
        const int cur = jt & 1;
        if (tid == 0 && jt + 1 < nTiles) {
            const unsigned bytes = (unsigned)tile_atoms(jt + 1) * 16u;
            mbar_expect_tx(&t.mbar[cur ^ 1], bytes);
            tma_bulk_g2s(t.sP + (cur ^ 1) * SURF_TILE, Pf + (size_t)(jt + 1) * SURF_TILE, bytes, &t.mbar[cur ^ 1]);
        }
        if (cur == 0) { mbar_wait(&t.mbar[0], phase0); phase0 ^= 1u; }
        else          { mbar_wait(&t.mbar[1], phase1); phase1 ^= 1u; }
        const float4 *tile = t.sP + cur * SURF_TILE;
        const int nb = tile_atoms(jt);
        for (int j = 0; j < nb; ++j) body(jt * SURF_TILE + j, tile[j]);
        __syncthreads();
    }
}

// water [nMol*nWAtoms][nF][3] (device, reference layout) shifted in place for frames [f0, f0+nFc).
// grid (ceil(nMol/SURF_THREADS), nFc).
__global__ void __launch_bounds__(SURF_THREADS)
water_nearest_shift_kernel(float *__restrict__ water, int nMol, int nWAtoms, int nF, int f0,
                           const float4 *__restrict__ P, int nP, int nPPad, const float *__restrict__ cell)
{
    __shared__ __align__(128) float4 sP[2 * SURF_TILE];
    __shared__ uint64_t mbar[2];
    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); mbar_fence_init(); }
    __syncthreads();
    unsigned phase0 = 0, phase1 = 0;
    const SurfTiles tiles{sP, mbar};

    const int fl = blockIdx.y, f = f0 + fl;
    const int m = blockIdx.x * SURF_THREADS + tid;
    const bool live = m < nMol;
    const float cx = cell[3 * (size_t)f], cy = cell[3 * (size_t)f + 1], cz = cell[3 * (size_t)f + 2];
    const float icx = __frcp_rn(cx), icy = __frcp_rn(cy), icz = __frcp_rn(cz);
    float *wp = water + 3 * ((size_t)nWAtoms * nF * (live ? m : 0) + f);
    const float wx = live ? wp[0] : 0.f, wy = live ? wp[1] : 0.f, wz = live ? wp[2] : 0.f;

    float best = 1000.0f, sx = 0.f, sy = 0.f, sz = 0.f;      // setWaterDistPBC.cpp:13-16, per molecule
    for_each_protein_atom(P + (size_t)fl * nPPad, nP, tiles, phase0, phase1, [&](int, const float4 &p) {
        float dx, dy, dz, px, py, pz;
        strict_wrap_with_shift(__fsub_rn(wx, p.x), cx, icx, dx, px);       // :42-53, water - protein
        strict_wrap_with_shift(__fsub_rn(wy, p.y), cy, icy, dy, py);
        strict_wrap_with_shift(__fsub_rn(wz, p.z), cz, icz, dz, pz);
        const float dist = __fsqrt_rn(strict_norm2(dx, dy, dz));           // :55
        if (dist < best) { best = dist; sx = px; sy = py; sz = pz; }       // :57-63 (first minimum wins)
    });
    if (live) {
        for (int k = 0; k < nWAtoms; ++k) {                                  // :67-72
            float *q = wp + 3 * (size_t)k * nF;
            q[0] = __fsub_rn(q[0], sx);
            q[1] = __fsub_rn(q[1], sy);
            q[2] = __fsub_rn(q[2], sz);
        }
    }
}

// waterO [nO][nF][3] overwritten in place for frames [f0, f0+nFc): (cos, flag, nearest distance).
// grid (ceil(nO/SURF_THREADS), nFc).
__global__ void __launch_bounds__(SURF_THREADS)
water_orient_kernel(float *__restrict__ waterO, const float *__restrict__ watVec, int nO, int nF, int f0,
                    const float4 *__restrict__ P, int nP, int nPPad, const float *__restrict__ cell,
                    float minR, float maxR, int maxN)
{
    __shared__ __align__(128) float4 sP[2 * SURF_TILE];
    __shared__ uint64_t mbar[2];
    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); mbar_fence_init(); }
    __syncthreads();
    unsigned phase0 = 0, phase1 = 0;
    const SurfTiles tiles{sP, mbar};

    const int fl = blockIdx.y, f = f0 + fl;
    const int w = blockIdx.x * SURF_THREADS + tid;
    const bool live = w < nO;
    const float cx = cell[3 * (size_t)f], cy = cell[3 * (size_t)f + 1], cz = cell[3 * (size_t)f + 2];
    const float icx = __frcp_rn(cx), icy = __frcp_rn(cy), icz = __frcp_rn(cz);
    float *wo = waterO + 3 * ((size_t)nF * (live ? w : 0) + f);
    const float *wv = watVec + 3 * ((size_t)nF * (live ? w : 0) + f);
    const float wx = live ? wo[0] : 0.f, wy = live ? wo[1] : 0.f, wz = live ? wo[2] : 0.f;

    float cl[5 * SURF_MAXN];                                  // (atom, dist^2, dx, dy, dz) per slot
    for (int i = 0; i < 5 * maxN; ++i) cl[i] = 1000.0f;       // waterOrientAtSurface.cpp:27-28
    float flag = 0.0f;                                        // :43

    for_each_protein_atom(P + (size_t)fl * nPPad, nP, tiles, phase0, phase1, [&](int pAtom, const float4 &p) {
        float dx, dy, dz, px, py, pz;
        strict_wrap_with_shift(__fsub_rn(p.x, wx), cx, icx, dx, px);       // :47-54, protein - water
        strict_wrap_with_shift(__fsub_rn(p.y, wy), cy, icy, dy, py);
        strict_wrap_with_shift(__fsub_rn(p.z, wz), cz, icz, dz, pz);
        const float dist = __fsqrt_rn(strict_norm2(dx, dy, dz));           // :56
        if (dist <= maxR && dist >= minR) {                                 // :58
            flag = 1.0f;
            for (int i = 0; i < maxN; ++i) {                                // :62
                if (dist < __fsqrt_rn(cl[5 * i + 1])) {                     // :64
                    // :65-72 literally: the copy runs upwards, so every later slot receives old slot i
                    for (int k = i + 1; k < maxN; ++k) {
                        cl[5 * k] = cl[5 * (k - 1)]; cl[5 * k + 1] = cl[5 * (k - 1) + 1];
                        cl[5 * k + 2] = cl[5 * (k - 1) + 2]; cl[5 * k + 3] = cl[5 * (k - 1) + 3];
                        cl[5 * k + 4] = cl[5 * (k - 1) + 4];
                    }
                    cl[5 * i] = (float)pAtom;
                    cl[5 * i + 1] = __fmul_rn(dist, dist);
                    cl[5 * i + 2] = dx; cl[5 * i + 3] = dy; cl[5 * i + 4] = dz;
                    break;
                }
            }
        }
    });
    if (live) {
        float n0 = 0.f, n1 = 0.f, n2 = 0.f;                                  // :91-98
        for (int i = 0; i < maxN; ++i) {
            n0 = __fadd_rn(n0, __fdiv_rn(cl[5 * i + 2], cl[5 * i + 1]));
            n1 = __fadd_rn(n1, __fdiv_rn(cl[5 * i + 3], cl[5 * i + 1]));
            n2 = __fadd_rn(n2, __fdiv_rn(cl[5 * i + 4], cl[5 * i + 1]));
        }
        const float vx = wv[0], vy = wv[1], vz = wv[2];
        float cosA = __fadd_rn(__fadd_rn(__fmul_rn(vx, n0), __fmul_rn(vy, n1)), __fmul_rn(vz, n2));      // :101
        cosA = __fdiv_rn(cosA, __fsqrt_rn(strict_norm2(vx, vy, vz)));                                    // :102
        cosA = __fdiv_rn(cosA, __fsqrt_rn(strict_norm2(n0, n1, n2)));                                    // :103
        wo[0] = cosA;                                                        // :106
        wo[1] = flag;                                                        // :43,60
        wo[2] = __fsqrt_rn(cl[1]);                                           // :107
    }
}

// ------------------------------------------------------------------------------------------
// hydrogen-bond autocorrelation (lib/openmp/src/getHydrogenBonds.cpp:9-91)
//
// Only pairs (acceptor row, donor/hydrogen col > row) bonded at dt = 0 contribute (:78-79), so the
// work splits into
//   hb_frame0_kernel  every pair at frame 0 with the strict criterion -> compact list of bonded pairs
//   hb_walk_kernel    one warp per bonded pair, lanes = 32 consecutive frames (coalesced rows of the
//                     reference layout); intermittent: count every bonded frame; continuous: count
//                     the leading run of bonded frames and stop at the first break (:81-82)
// Counts are integers (the reference adds 1.0f per event into a float under `omp critical`).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool hb_bonded(const float *__restrict__ a, const float *__restrict__ h,
                                          const float *__restrict__ d, float cx, float cy, float cz,
                                          float maxR, float cosAngle)
{
    float hx, hy, hz, sx, sy, sz;
    strict_wrap_with_shift(__fsub_rn(h[0], a[0]), cx, __frcp_rn(cx), hx, sx);     // :37-49, h - acc, wrapped
    strict_wrap_with_shift(__fsub_rn(h[1], a[1]), cy, __frcp_rn(cy), hy, sy);
    strict_wrap_with_shift(__fsub_rn(h[2], a[2]), cz, __frcp_rn(cz), hz, sz);
    const float ax = __fsub_rn(a[0], d[0]), ay = __fsub_rn(a[1], d[1]), az = __fsub_rn(a[2], d[2]);   // :52-57, NOT wrapped
    const float hh = strict_norm2(hx, hy, hz);
    const float dist = __fsqrt_rn(hh);                                             // :60
    const float dot = __fadd_rn(__fadd_rn(__fmul_rn(hx, ax), __fmul_rn(hy, ay)), __fmul_rn(hz, az));   // :63
    const float den = __fmul_rn(__fsqrt_rn(hh), __fsqrt_rn(strict_norm2(ax, ay, az)));                 // :64-65
    const float angle = __fdiv_rn(dot, den);
    return dist < maxR && angle < cosAngle;                                        // :67
}

// grid (ceil(nD/256), min(nA, 4096)), block 256.  pairs[2*k], pairs[2*k+1] = (row, col).
__global__ void __launch_bounds__(256)
hb_frame0_kernel(const float *__restrict__ acc, int nA, const float *__restrict__ don, int nD,
                 const float *__restrict__ hyd, const float *__restrict__ cell, int nF,
                 float maxR, float cosAngle, int *__restrict__ pairs, unsigned capacity,
                 unsigned *__restrict__ nPairs)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= nD) return;
    const float cx = cell[0], cy = cell[1], cz = cell[2];
    const float *h = hyd + (size_t)col * nF * 3;
    const float *d = don + (size_t)col * nF * 3;
    for (int row = blockIdx.y; row < nA && row < col; row += gridDim.y) {
        if (hb_bonded(acc + (size_t)row * nF * 3, h, d, cx, cy, cz, maxR, cosAngle)) {
            const unsigned k = atomicAdd(nPairs, 1u);
            if (k < capacity) { pairs[2 * k] = row; pairs[2 * k + 1] = col; }
        }
    }
}

// one warp per bonded pair; grid ceil(nPairs/8), block 256
__global__ void __launch_bounds__(256)
hb_walk_kernel(const float *__restrict__ acc, const float *__restrict__ don, const float *__restrict__ hyd,
               const float *__restrict__ cell, int nF, float maxR, float cosAngle, int continuous,
               const int *__restrict__ pairs, unsigned nPairs, unsigned long long *__restrict__ counts)
{
    const unsigned pidx = blockIdx.x * 8u + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (pidx >= nPairs) return;
    const int row = pairs[2 * pidx], col = pairs[2 * pidx + 1];
    const float *a = acc + (size_t)row * nF * 3;
    const float *h = hyd + (size_t)col * nF * 3;
    const float *d = don + (size_t)col * nF * 3;
    for (int dt0 = 0; dt0 < nF; dt0 += 32) {
        const int dt = dt0 + lane;
        bool bonded = false;
        if (dt < nF)
            bonded = hb_bonded(a + 3 * dt, h + 3 * dt, d + 3 * dt, cell[3 * dt], cell[3 * dt + 1], cell[3 * dt + 2],
                               maxR, cosAngle);
        const unsigned mask = __ballot_sync(NDA_FULL_MASK, bonded);
        const unsigned valid = (nF - dt0 >= 32) ? NDA_FULL_MASK : ((1u << (nF - dt0)) - 1u);
        if (!continuous) {
            if (bonded) atomicAdd(&counts[dt], 1ull);                              // :69-73, notBroken stays 1
        } else {
            const unsigned broken = ~mask & valid;
            const int firstBreak = broken ? (__ffs(broken) - 1) : 32;
            if (bonded && lane < firstBreak) atomicAdd(&counts[dt], 1ull);
            if (broken) break;                                                      // :81-82: never counted again
        }
    }
}
