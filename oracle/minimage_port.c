/*
 * oracle/minimage_port.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, strict IEEE float32: build with
 * -fno-fast-math -ffp-contract=off) of the three minimum-image pair loops of
 * kpounot/NAMDAnalyzer's OpenMP build.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library;
 * the product (namdanalyzer_b200) never does.
 *
 * Parity pin: checked bit-for-bit against the reference's own sources compiled
 * where they lie (oracle/_ref/libref_strict.so, see oracle/Makefile and
 * tests/test_oracle.py) and against the golden vectors those produced
 * (tests/golden/).
 *
 * Each function cites the reference lines it follows; paths are relative to
 * /root/reference/package/NAMDAnalyzer/lib/openmp/src/.
 *
 * Deliberate differences from the reference text:
 *   B-3  getWithin.cpp:55 indexes outSel[atomId*nbrFrames + frame] (out of
 *        bounds for nbrFrames > 1).  Both branches it guards write the same
 *        value, so the result does not depend on it; the port drops the read
 *        and therefore handles any number of frames.
 *   B-4  getDistances.cpp:57 assigns (last frame / nF); the documented and
 *        legacy-CUDA behaviour is the mean over frames.  The port returns the
 *        float64 SUM over frames of the per-frame float32 distance so the
 *        caller can form either.
 *   B-8  getRadialNbrDensity.cpp:44 accumulates 1.0/nF in float32.  The port
 *        counts in uint64 (exactly what the reference produces when driven one
 *        frame per call, where the increment is 1.0).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* d - c*roundf(d/c): division, round-half-away, product and difference each
 * rounded once to binary32 (getWithin.cpp:63, getDistances.cpp:52-54,
 * getRadialNbrDensity.cpp:34-36). */
static inline float wrap1(float d, float c)
{
    float q = d / c;
    float n = roundf(q);
    float p = c * n;
    return d - p;
}

/* ((x*x + y*y) + z*z), every operation rounded (getWithin.cpp:73,
 * getDistances.cpp:56, getRadialNbrDensity.cpp:38). */
static inline float norm2(float x, float y, float z)
{
    float xx = x * x;
    float yy = y * y;
    float zz = z * z;
    float s = xx + yy;
    return s + zz;
}

/* getWithin.cpp:9-90.  allAtoms [nAtoms][nFrames][3], cellDims [nFrames][3],
 * out int32 [nAtoms][nFrames] (only ever set to 1, never cleared).
 * literal_early_out != 0 keeps the per-axis tests of lines 64-71, which compare
 * a signed component against distance squared; 0 gives the
 * pure predicate r2 <= distance^2. */
void port_get_within(const float *allAtoms, int nAtoms, int nFrames,
                     const int *refSel, int refSize,
                     const int *outSel, int outSelSize,
                     int *out, const float *cellDims, float distance,
                     int literal_early_out)
{
    (void)nAtoms;
    const float squaredDist = distance * distance;          /* :29 */
    #pragma omp parallel for schedule(static)
    for (int frame = 0; frame < nFrames; ++frame) {        /* :31 */
        const float cx = cellDims[3 * frame];
        const float cy = cellDims[3 * frame + 1];
        const float cz = cellDims[3 * frame + 2];
        for (int r = 0; r < refSize; ++r) {                 /* :37 */
            const int refIdx = refSel[r];
            const float *ref = allAtoms + 3 * ((size_t)nFrames * refIdx + frame);
            const float sx = ref[0], sy = ref[1], sz = ref[2];
            for (int a = 0; a < outSelSize; ++a) {          /* :46 */
                const int atomIdx = outSel[a];
                const float *at = allAtoms + 3 * ((size_t)nFrames * atomIdx + frame);
                /* :55-86 -- with out==1 and refIdx==atomIdx the else-branch
                 * rewrites the 1 that is already there; otherwise compute. */
                float dx = at[0] - sx;                      /* :58-60 atom - ref */
                float dy = at[1] - sy;
                float dz = at[2] - sz;
                dx = wrap1(dx, cx);                         /* :63 */
                if (literal_early_out && dx > squaredDist) continue;   /* :64-65 */
                dy = wrap1(dy, cy);
                if (literal_early_out && dy > squaredDist) continue;   /* :67-68 */
                dz = wrap1(dz, cz);
                if (literal_early_out && dz > squaredDist) continue;   /* :70-71 */
                if (norm2(dx, dy, dz) <= squaredDist)       /* :73-75 */
                    out[(size_t)atomIdx * nFrames + frame] = 1;        /* :77 */
            }
        }
    }
}

/* getDistances.cpp:9-63.  sel1 [n1][nF][3], sel2 [n2][nF][3].
 * sum64[i*n2+j] += sqrtf(r2) for every frame; with sameSel only j > i is
 * visited (:44) and the rest of sum64 is left untouched.
 * If last32 != NULL it receives the per-frame float32 distance of the LAST
 * frame (what the reference's assignment leaves behind, before its /nF). */
void port_get_distances(const float *sel1, int n1, const float *sel2, int n2,
                        double *sum64, float *last32,
                        const float *cellDims, int nFrames, int sameSel)
{
    #pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < n1; ++i) {                          /* :35 */
        for (int frame = 0; frame < nFrames; ++frame) {    /* :26 */
            const float cx = cellDims[3 * frame];
            const float cy = cellDims[3 * frame + 1];
            const float cz = cellDims[3 * frame + 2];
            const float *a = sel1 + 3 * ((size_t)i * nFrames + frame);
            for (int j = (sameSel == 1 ? i + 1 : 0); j < n2; ++j) {   /* :44 */
                const float *b = sel2 + 3 * ((size_t)j * nFrames + frame);
                float dx = b[0] - a[0];                     /* :47-49 sel2 - sel1 */
                float dy = b[1] - a[1];
                float dz = b[2] - a[2];
                dx = wrap1(dx, cx);                         /* :52-54 */
                dy = wrap1(dy, cy);
                dz = wrap1(dz, cz);
                float dist = sqrtf(norm2(dx, dy, dz));      /* :56 float overload */
                sum64[(size_t)i * n2 + j] += (double)dist;
                if (last32) last32[(size_t)i * n2 + j] = dist;
            }
        }
    }
}

/* getRadialNbrDensity.cpp:8-52.  counts[g*outSize + idx] += 1 for every
 * (row, col, frame) with 0 <= idx < outSize, idx = (int)(sqrtf(r2)/dr) (:38-41).
 * groupId == NULL means a single group (the reference call); otherwise
 * groupId[col] in [0, nGroups) selects the histogram row (one reference call
 * per residue, dataAnalysis/RadialDensity.py:84-103, batched).
 * maxR is accepted and ignored exactly as the reference ignores it (B-9). */
void port_get_rdf_counts(const float *sel1, int n1, const float *sel2, int n2,
                         uint64_t *counts, int outSize, const int *groupId, int nGroups,
                         const float *cellDims, int nFrames, int sameSel,
                         float maxR, float dr)
{
    (void)maxR;
    const size_t H = (size_t)outSize * (size_t)(groupId ? nGroups : 1);
    #pragma omp parallel
    {
        uint64_t *local = (uint64_t *)calloc(H ? H : 1, sizeof(uint64_t));
        #pragma omp for schedule(dynamic, 8) collapse(2)
        for (int frame = 0; frame < nFrames; ++frame) {    /* :12 */
            for (int row = 0; row < n1; ++row) {            /* :20 */
                const float cx = cellDims[3 * frame];
                const float cy = cellDims[3 * frame + 1];
                const float cz = cellDims[3 * frame + 2];
                const float *a = sel1 + 3 * ((size_t)row * nFrames + frame);
                for (int col = (sameSel == 1 ? row + 1 : 0); col < n2; ++col) {  /* :27 */
                    const float *b = sel2 + 3 * ((size_t)col * nFrames + frame);
                    float dx = a[0] - b[0];                 /* :29-31 sel1 - sel2 */
                    float dy = a[1] - b[1];
                    float dz = a[2] - b[2];
                    dx = wrap1(dx, cx);                     /* :34-36 */
                    dy = wrap1(dy, cy);
                    dz = wrap1(dz, cz);
                    float dist = sqrtf(norm2(dx, dy, dz));  /* :38 */
                    float t = dist / dr;                    /* :40 float / float */
                    /* (int)t is undefined in C for NaN / out-of-range; x86
                     * cvttss2si returns INT_MIN there, which fails idx >= 0. */
                    if (!(t > -1.0f && t < 2147483648.0f)) continue;
                    int idx = (int)t;
                    if (idx >= 0 && idx < outSize) {        /* :41 */
                        size_t g = groupId ? (size_t)groupId[col] : 0;
                        local[g * outSize + idx] += 1;      /* :44 with nF = 1 */
                    }
                }
            }
        }
        #pragma omp critical
        for (size_t k = 0; k < H; ++k) counts[k] += local[k];
        free(local);
    }
}

/* Per-pair probe used by the tests to enumerate pairs near an edge: returns the
 * strict float32 r2 for one pair in one frame with the sign convention of
 * getRadialNbrDensity.cpp:29-31 (a - b). */
float port_pair_r2(const float *a, const float *b, const float *cell)
{
    float dx = wrap1(a[0] - b[0], cell[0]);
    float dy = wrap1(a[1] - b[1], cell[1]);
    float dz = wrap1(a[2] - b[2], cell[2]);
    return norm2(dx, dy, dz);
}

int port_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * the two water-at-surface loops that reuse the minimum-image
 * inner loop.  Paths relative to lib/openmp/src/.
 * ------------------------------------------------------------------------------------------ */

/* setWaterDistPBC.cpp:9-77.  water [sizeW*nbrWAtoms][nF][3] is shifted IN PLACE, molecule by
 * molecule, by the periodic image vector c*roundf(d/c) of its nearest protein atom (distance
 * measured from the molecule's first atom, :33-35).
 *
 * reset_per_water == 0 is the literal OpenMP text: closestDist / closest_xyz are initialised
 * ONCE before all loops (:13-16), so they form a running minimum over every water and frame
 * processed so far and most molecules receive the stale shift of an earlier one.
 * reset_per_water == 1 re-initialises them for every molecule, which is what the legacy CUDA
 * build does (lib/cuda/src/setWaterDistPBC.cu:33) and what the docstring describes; the two
 * coincide when the reference is called with one molecule and one frame. */
void port_set_water_dist_pbc(float *water, int sizeW, const float *prot, int sizeP,
                             const float *cellDims, int nFrames, int nbrWAtoms, int reset_per_water)
{
    float closestDist = 1000, closest_x = 0, closest_y = 0, closest_z = 0;       /* :13-16 */
    for (int frame = 0; frame < nFrames; ++frame) {                              /* :22 */
        const float cx = cellDims[3 * frame], cy = cellDims[3 * frame + 1], cz = cellDims[3 * frame + 2];
        for (int w = 0; w < sizeW; ++w) {                                        /* :30 */
            if (reset_per_water) { closestDist = 1000; closest_x = closest_y = closest_z = 0; }
            const float *wp = water + 3 * ((size_t)nbrWAtoms * nFrames * w + frame);
            const float wx = wp[0], wy = wp[1], wz = wp[2];
            for (int p = 0; p < sizeP; ++p) {                                    /* :36 */
                const float *pp = prot + 3 * ((size_t)p * nFrames + frame);
                float dx = wx - pp[0], dy = wy - pp[1], dz = wz - pp[2];         /* :42-44 */
                const float px = cx * roundf(dx / cx);                           /* :47-49 */
                const float py = cy * roundf(dy / cy);
                const float pz = cz * roundf(dz / cz);
                dx -= px; dy -= py; dz -= pz;                                    /* :51-53 */
                const float dist = sqrtf(norm2(dx, dy, dz));                     /* :55 */
                if (dist < closestDist) {                                        /* :57-63 */
                    closestDist = dist;
                    closest_x = px; closest_y = py; closest_z = pz;
                }
            }
            for (int k = 0; k < nbrWAtoms; ++k) {                                /* :67-72 */
                float *q = water + 3 * ((size_t)nbrWAtoms * nFrames * w + frame + (size_t)k * nFrames);
                q[0] -= closest_x; q[1] -= closest_y; q[2] -= closest_z;
            }
        }
    }
}

/* waterOrientAtSurface.cpp:10-112.  waterO [sizeO][nF][3] is OVERWRITTEN in place:
 * [..][0] = cosine between watVec and the surface normal built from the maxN nearest protein
 * atoms with minR <= dist <= maxR, [..][1] = 1 if any such atom exists else 0,
 * [..][2] = sqrtf(closest[1]) (distance of the nearest one, sqrtf(1000) if none). */
void port_water_orient_at_surface(float *waterO, int sizeO, const float *watVec, const float *prot, int sizeP,
                                  const float *cellDims, int nFrames, float minR, float maxR, int maxN)
{
    #pragma omp parallel
    {
        float *closest = (float *)malloc(sizeof(float) * 5 * (size_t)(maxN > 0 ? maxN : 1));
        #pragma omp for collapse(2) schedule(static)
        for (int frame = 0; frame < nFrames; ++frame) {                          /* :19 */
            for (int w = 0; w < sizeO; ++w) {                                    /* :31 */
                const float cx = cellDims[3 * frame], cy = cellDims[3 * frame + 1], cz = cellDims[3 * frame + 2];
                for (int i = 0; i < 5 * maxN; ++i) closest[i] = 1000;            /* :27-28 */
                float *wo = waterO + 3 * ((size_t)nFrames * w + frame);
                const float *wv = watVec + 3 * ((size_t)nFrames * w + frame);
                const float wx = wo[0], wy = wo[1], wz = wo[2];
                const float vx = wv[0], vy = wv[1], vz = wv[2];
                wo[1] = 0;                                                       /* :43 */
                for (int p = 0; p < sizeP; ++p) {                                /* :45 */
                    const float *pp = prot + 3 * ((size_t)p * nFrames + frame);
                    float dx = pp[0] - wx, dy = pp[1] - wy, dz = pp[2] - wz;     /* :47-49 prot - water */
                    dx = wrap1(dx, cx); dy = wrap1(dy, cy); dz = wrap1(dz, cz);  /* :52-54 */
                    const float dist = sqrtf(norm2(dx, dy, dz));                 /* :56 */
                    if (dist <= maxR && dist >= minR) {                          /* :58 */
                        wo[1] = 1;
                        for (int i = 0; i < maxN; ++i) {                         /* :62 */
                            if (dist < sqrtf(closest[5 * i + 1])) {              /* :64 */
                                /* :65-72 LITERALLY: the copy runs upwards, so it is not a shift --
                                 * every slot after i ends up holding the OLD slot i */
                                for (int k = i + 1; k < maxN; ++k)
                                    memcpy(closest + 5 * k, closest + 5 * (k - 1), 5 * sizeof(float));
                                closest[5 * i] = (float)p;
                                closest[5 * i + 1] = dist * dist;
                                closest[5 * i + 2] = dx; closest[5 * i + 3] = dy; closest[5 * i + 4] = dz;
                                break;
                            }
                        }
                    }
                }
                float n0 = 0, n1 = 0, n2 = 0;                                    /* :91-98 */
                for (int i = 0; i < maxN; ++i) {
                    n0 += closest[5 * i + 2] / closest[5 * i + 1];
                    n1 += closest[5 * i + 3] / closest[5 * i + 1];
                    n2 += closest[5 * i + 4] / closest[5 * i + 1];
                }
                float cosA = vx * n0 + vy * n1 + vz * n2;                        /* :101-103 */
                cosA /= sqrtf(vx * vx + vy * vy + vz * vz);
                cosA /= sqrtf(n0 * n0 + n1 * n1 + n2 * n2);
                wo[0] = cosA;                                                    /* :106-107 */
                wo[2] = sqrtf(closest[1]);
            }
        }
        free(closest);
    }
}

/* ------------------------------------------------------------------------------------------
 * hydrogen-bond autocorrelation.
 * getHydrogenBonds.cpp:9-91 (getHBCorr).  acceptors [nA][nF][3], donors [nD][nF][3],
 * hydrogens [nH][nF][3] (hydrogen col belongs to donor col).  For every pair (row, col > row)
 * that is bonded at dt = 0 -- |h - acc| (minimum image) < maxR and the cosine between (h - acc)
 * and the UNWRAPPED (acc - donor) below cosf(minAngle) -- counts[dt] += 1 for every later frame
 * at which it is bonded (intermittent) or still unbroken (continuous).  The reference adds
 * 1.0f into a float per event (:73); the port counts in uint64.  maxTime, step and nbrTimeOri
 * are accepted and unused, as in the reference.
 * ------------------------------------------------------------------------------------------ */
void port_get_hb_corr(const float *acceptors, int nA, int nFrames, const float *donors, int nD,
                      const float *hydrogens, int nH, const float *cellDims, uint64_t *counts,
                      float maxR, float minAngle, int continuous)
{
    (void)nH;
    const float cosAngle = cosf(minAngle);                                       /* :16 */
    #pragma omp parallel
    {
        uint64_t *local = (uint64_t *)calloc((size_t)(nFrames > 0 ? nFrames : 1), sizeof(uint64_t));
        #pragma omp for schedule(dynamic, 4)
        for (int row = 0; row < nA; ++row) {                                     /* :23 */
            for (int col = row + 1; col < nD; ++col) {                           /* :25 */
                int notBroken = 1;
                for (int dt = 0; dt < nFrames; ++dt) {                           /* :30 */
                    const float cx = cellDims[3 * dt], cy = cellDims[3 * dt + 1], cz = cellDims[3 * dt + 2];
                    const float *a = acceptors + 3 * ((size_t)nFrames * row + dt);
                    const float *h = hydrogens + 3 * ((size_t)nFrames * col + dt);
                    const float *d = donors + 3 * ((size_t)nFrames * col + dt);
                    float hx = h[0] - a[0], hy = h[1] - a[1], hz = h[2] - a[2]; /* :37-44 */
                    hx = wrap1(hx, cx); hy = wrap1(hy, cy); hz = wrap1(hz, cz);  /* :47-49 */
                    const float ax = a[0] - d[0], ay = a[1] - d[1], az = a[2] - d[2];   /* :52-57, not wrapped */
                    const float hh = norm2(hx, hy, hz);
                    const float dist = sqrtf(hh);                                /* :60 */
                    float angle = (hx * ax + hy * ay) + hz * az;                 /* :63 */
                    angle /= (sqrtf(hh) * sqrtf(norm2(ax, ay, az)));            /* :64-65 */
                    if (dist < maxR && angle < cosAngle) {                       /* :67 */
                        local[dt] += (uint64_t)notBroken;                        /* :69-73 (t0 = 1 here) */
                    } else {
                        if (dt == 0) break;                                      /* :78-79 */
                        if (continuous == 1) notBroken = 0;                      /* :81-82 */
                    }
                }
            }
        }
        #pragma omp critical
        for (int dt = 0; dt < nFrames; ++dt) counts[dt] += local[dt];
        free(local);
    }
}
