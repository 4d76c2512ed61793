/*
 * nda_b200.h -- C ABI of libnda_b200.so: the minimum-image pair path of
 * kpounot/NAMDAnalyzer (getWithin / getDistances / getRadialNbrDensity) as
 * hand-written sm_100a CUDA.  Plain pointers and sizes only; no torch, no C++
 * types.  Every function returns 0 on success and a non-zero code on failure;
 * nda_last_error() then holds a message (thread-local).  Nothing here ever
 * calls exit() (the reference's CUDA build does: lib/cuda/src/getWithin.cu:9-17).
 *
 * Reference interface replaced (paths relative to
 * /root/reference/package/NAMDAnalyzer/):
 *   lib/libFunc.h:37-38   getDistances          -> nda_get_distances
 *   lib/libFunc.h:41-43   getRadialNbrDensity   -> nda_get_radial_nbr_density
 *   lib/libFunc.h:46-48   getWithin             -> nda_get_within
 *   lib/libFunc.h:57      getParallelBackend    -> nda_get_parallel_backend
 * which the reference binds from Cython at lib/libFunc.pyx:118-127, 131-140,
 * 176-186, 356-358.  Argument order and meaning are the reference's, so the two
 * headers can be diffed.
 *
 * Layouts (all C-contiguous, the reference's own):
 *   coordinates  float32 [nAtoms][nFrames][3]   (dataParsers/dcdReader.py:127,175)
 *   cellDims     float32 [nFrames][3]           (dataParsers/dcdCell.py:113-114)
 *
 * "host" entry points take host pointers (pageable or pinned) and do the
 * host<->device copies themselves.  "dev" entry points take device pointers in
 * the same layouts, enqueue on the given CUDA stream (a cudaStream_t passed as
 * void*, NULL = the library's own stream) and do not synchronise, with one exception: the
 * within / RDF calls read the cell edges of the frame range back (12 bytes per frame, one
 * stream synchronisation) to choose the filter; the read-back is cached per (pointer, frame range)
 * and refreshed every 64th call.  All nda_dev_* calls of a device share one set of scratch
 * buffers: the library orders consecutive calls with an event, also across different caller streams.
 * One call at a time per DEVICE (a per-device lock); calls on different devices run concurrently.
 */
#ifndef NDA_B200_H
#define NDA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status ------------------------------------------------------------------- */
#define NDA_OK              0
#define NDA_ERR_ARG         1   /* bad argument (null pointer, negative size, ...)   */
#define NDA_ERR_CUDA        2   /* a CUDA runtime call or kernel failed             */
#define NDA_ERR_NO_DEVICE   3   /* no usable sm_100 device                          */

const char *nda_last_error(void);
/* lib/libFunc.h:57 -- 2 = "CUDA build" so NAMDDCD.getWithin always takes the native
 * branch and never the scipy kd-tree (dataParsers/dcdParser.py:181). */
int  nda_get_parallel_backend(void);
int  nda_device_count(void);
/* Devices of the CALLING THREAD.  nda_set_device: one device (also the one the nda_dev_* entry points
 * use).  nda_set_devices: the host entry points (nda_get_within*, nda_get_radial_nbr_*, nda_get_distances*)
 * split their frame range [floor(g nF / G), floor((g + 1) nF / G)) over the G devices -- the frame is the
 * outermost loop of every reference function (lib/openmp/src/getWithin.cpp:31, getDistances.cpp:26,
 * getRadialNbrDensity.cpp:12) -- one host thread and one context per device, every device uploading
 * only its own frame columns over its own link; histograms (u64) and distance sums (f64) are merged by
 * ONE ncclAllReduce per call on the compute streams (communicators from ncclCommInitAll; NCCL is
 * dlopen()ed), within needs no collective.  Default: NDA_DEVICES ("0,1,2" or "all") if set, else the
 * current CUDA device.  nda_get_devices returns the count (and up to `capacity` ids). */
int  nda_set_device(int device);
int  nda_set_devices(const int *devices, int n);
int  nda_get_devices(int *devices, int capacity);
int  nda_get_device(void);
/* how the last multi-device host call merged its partial results: 0 no merge (one device, or within),
 * 1 ncclAllReduce, 2 summed on the host (NCCL not loadable, or NDA_COLLECTIVE=host) */
int  nda_last_collective(void);
/* Tuning / debugging switches.  The environment variables of the same names (NDA_FILTER,
 * NDA_TC_RANGE_FACTOR, NDA_TC_MAX_FRACTION, NDA_TC_ALLOW_PADDING, NDA_NO_CELL_CACHE, NDA_STRICT_SLOW,
 * NDA_TC_NO_SURE, NDA_CTAS_PER_SM, NDA_DEBUG, NDA_HOST_THREADS, NDA_UPLOAD_CHUNKS, NDA_WITHIN_CHUNKS,
 * NDA_NO_TAPER, NDA_TAPER_LEVELS, NDA_COPY_STREAMS, NDA_D2H_COPY_ENGINE, NDA_COLLECTIVE=nccl|host,
 * NDA_SUBPASS_FRAMES, NDA_COMPAT_CUDA, NDA_NVTX, NDA_HEAD_TAPER, NDA_SIDE_TEAM, NDA_DENSE_FRACTION) are read ONCE, when the library is first used; afterwards
 * only this call changes a value.  value NULL or "" restores the default.  No switch changes a result,
 * except NDA_COMPAT_CUDA=1: nda_get_within then also marks every refSel atom in every frame, as the
 * reference's legacy CUDA build does (lib/cuda/src/getWithin.cu:27-28). */
int  nda_set_option(const char *name, const char *value);
/* frees every device buffer, pinned staging area, stream, event and NCCL communicator the library
 * holds (they are kept between calls otherwise); the next call builds them again */
int  nda_shutdown(void);
/* kernels launched by this library in this process since the last reset */
uint64_t nda_launch_count(void);
void     nda_launch_count_reset(void);
/* 0 (default): literal getWithin.cpp:64-71 per-axis early-outs; 1: pure r2 <= d^2 */
int  nda_set_within_pure_predicate(int on);
/* filter the last within / RDF call ran: 100 = tensor-core filter (csrc/nda_tc.cuh), 0 = rint,
 * 1 = fold, 2..5 = packed-FP32 fold2 variants (csrc/nda_device.cuh); -1 before the first call.
 * NDA_FILTER=tc|rint|fold|fold2|fold2a0..3 selects; tc (default) falls back to fold2 when a frame
 * of the call is not eligible (cut-off too close to the cell size, no unit cell, dense ranges) */
int  nda_last_filter(void);
/* debugging aid (trace builds of tools/build_variants.py only): with NDA_TC_TRACE=1 the last tensor-core
 * within call recorded clock64 stamps of CTA 0's pipeline (256 steps x 32 stamps, csrc/nda_tc.cuh: tc_stamp) */
int  nda_debug_tc_trace(long long *out);
/* parity support: ONE production tcgen05.mma (M = 128, N = 256, K = 16, fp16 accumulators; the instruction
 * descriptor, operand layout and tcgen05.ld.pack::16b read-back of the pair kernels) on caller-supplied
 * operands.  A uint16 [128][16], B uint16 [256][16]: fp16 bit patterns, row-major (host pointers).
 * out uint32 [128][128]: out[row][32 * cq + n] = register n of the epilogue warp that owns columns
 * 64 cq .. 64 cq + 63 of that row.  tests/test_gpu_tensor_filter.py pins with it that an accumulator is the
 * correctly rounded exact sum and that register n holds columns 2n (low half) and 2n + 1 (high half). */
int  nda_debug_tc_probe(const uint16_t *A, const uint16_t *B, uint32_t *out);

/* ---- drop-in host entry points (reference signatures) --------------------------- */

/* lib/openmp/src/getWithin.cpp:9-90.  out int32 [nbrAtoms][nbrFrames]: entries of
 * outSel atoms within `distance` of any refSel atom are set to 1; nothing is cleared. */
int nda_get_within(const float *allAtoms, int nbrAtoms, int nbrFrames,
                   const int *refSel, int refSize, const int *outSel, int outSelSize,
                   int *out, const float *cellDims, float distance);

/* lib/openmp/src/getDistances.cpp:9-63.  out float32 [sel1_size][sel2_size] = mean over
 * frames of the minimum-image distance (the documented behaviour,
 * dataParsers/dcdParser.py:103-105, and what lib/cuda/src/getDistances.cu:49 computes).
 * sameSel != 0: sel2 is sel1; the matrix is filled symmetrically, diagonal 0. */
int nda_get_distances(const float *sel1, int sel1_size, const float *sel2, int sel2_size,
                      float *out, const float *cellDims, int nbrFrames, int sameSel);

/* lib/openmp/src/getRadialNbrDensity.cpp:8-52.  count[idx] / nbrFrames, idx = (int)(dist/dr), is
 * ADDED to out float32 [outSize], as the reference accumulates (:40-41; every stock caller passes
 * zeros); sameSel != 0 counts each unordered pair once (col > row).  maxR is accepted and ignored,
 * as in the reference. */
int nda_get_radial_nbr_density(const float *sel1, int sel1_size, const float *sel2, int sel2_size,
                               float *out, int outSize, const float *cellDims, int nbrFrames,
                               int sameSel, float maxR, float dr);

/* ---- extended host entry points -------------------------------------------------- */

/* Exact integer pair counts and/or the batched per-group histogram that
 * dataAnalysis/RadialDensity.py:84-103 obtains with one reference call per residue.
 * groupId int32 [sel2_size] in [0, nGroups) (NULL: one group).  counts uint64
 * [nGroups][outSize] and/or out float32 [nGroups][outSize] (either may be NULL) are
 * OVERWRITTEN.  Only frames [frameBegin, frameEnd) of the nbrFrames-frame arrays are
 * processed (the multi-GPU shard); out is still normalised by nbrFrames. */
int nda_get_radial_nbr_counts(const float *sel1, int sel1_size, const float *sel2, int sel2_size,
                              const int *groupId, int nGroups,
                              uint64_t *counts, float *out, int outSize,
                              const float *cellDims, int nbrFrames, int frameBegin, int frameEnd,
                              int sameSel, float dr);

/* Sum over frames [frameBegin, frameEnd) of the per-frame float32 distance, in float64
 * [sel1_size][sel2_size] (OVERWRITTEN); sameSel fills both triangles. */
int nda_get_distances_sum(const float *sel1, int sel1_size, const float *sel2, int sel2_size,
                          double *sum, const float *cellDims, int nbrFrames,
                          int frameBegin, int frameEnd, int sameSel);

/* The within bitmask itself: bits uint32 [frameEnd-frameBegin][ceil(outSelSize/32)]
 * (OVERWRITTEN), bit (a & 31) of word a >> 5 = outSel[a] is within.  out may be NULL. */
int nda_get_within_bits(const float *allAtoms, int nbrAtoms, int nbrFrames,
                        const int *refSel, int refSize, const int *outSel, int outSelSize,
                        uint32_t *bits, int *out, const float *cellDims, float distance,
                        int frameBegin, int frameEnd);

/* ---- the two water-at-surface loops that reuse the minimum-image inner loop ---- */

/* lib/libFunc.h:53-54 setWaterDistPBC, lib/openmp/src/setWaterDistPBC.cpp:9-77.  water float32
 * [sizeW*nbrWAtoms][nbrFrames][3] is shifted IN PLACE, molecule by molecule, by the periodic image
 * vector c*roundf(d/c) of the protein atom nearest to the molecule's first atom.  The minimum is
 * taken PER MOLECULE (legacy CUDA build, lib/cuda/src/setWaterDistPBC.cu:33, and the docstring);
 * the OpenMP text carries closestDist over from one molecule to the next. */
int nda_set_water_dist_pbc(float *water, int sizeW, const float *prot, int sizeP,
                           const float *cellDims, int nbrFrames, int nbrWAtoms);

/* lib/libFunc.h:50-51 waterOrientAtSurface, lib/openmp/src/waterOrientAtSurface.cpp:10-112.
 * waterO float32 [sizeO][nbrFrames][3] is OVERWRITTEN in place with (cosine of watVec against the
 * surface normal built from the maxN nearest protein atoms with minR <= dist <= maxR, 0/1 flag,
 * distance of the nearest one); `out` is accepted and ignored like in the reference. 1 <= maxN <= 16. */
int nda_water_orient_at_surface(float *waterO, int sizeO, const float *watVec, const float *prot, int sizeP,
                                float *out, const float *cellDims, int nbrFrames,
                                float minR, float maxR, int maxN);

/* ---- hydrogen-bond autocorrelation ------------------------------------ */

/* lib/libFunc.h:13-19 getHBCorr, lib/openmp/src/getHydrogenBonds.cpp:9-91.  acceptors
 * [size_acceptors][nbrFrames][3], donors / hydrogens [size_donors][nbrFrames][3] (hydrogen col is
 * bound to donor col).  out float32 [nbrFrames] is ACCUMULATED into, as in the reference:
 * out[dt] += number of pairs (row, col > row) bonded at dt = 0 that are bonded (continuous == 0)
 * or still unbroken (continuous == 1) at dt.  maxTime, step, nbrTimeOri are accepted and unused,
 * as in the reference.  The reference adds 1.0f per event; this adds the integer count once
 * (identical while a call's counts stay below 2^24). */
int nda_get_hb_corr(const float *acceptors, int size_acceptors, int nbrFrames,
                    const float *donors, int size_donors, const float *hydrogens, int size_hydrogens,
                    const float *cellDims, float *out, int maxTime, int step, int nbrTimeOri,
                    float maxR, float minAngle, int continuous);
/* the exact integer counts, uint64 [nbrFrames], OVERWRITTEN */
int nda_get_hb_counts(const float *acceptors, int size_acceptors, int nbrFrames,
                      const float *donors, int size_donors, const float *hydrogens, int size_hydrogens,
                      const float *cellDims, uint64_t *counts, float maxR, float minAngle, int continuous);

/* ---- DCD ingest (host I/O feeding the path) ---------------------------- */

/* lib/libFunc.h:21-28 getDCDCoor / lib/common/src/dcdReader.cpp:28-98.  Reads the records of the
 * requested frames / axes and gathers the selected atoms into outArr float32
 * [selAtomsSize][nbrFrames][nbrDims].  Returns the reference's codes: 0, -1 (cannot open),
 * -2 (bad index / seek), -3 (short read).  Frames are read by several host threads (pread). */
int nda_dcd_get_coor(const char *fileName, const int *frames, int nbrFrames, int nbrAtoms,
                     const int *selAtoms, int selAtomsSize, const int *dims, int nbrDims, int cell,
                     const long long *startPos, float *outArr, char byteorder);
/* lib/libFunc.h:30-31 getDCDCell / lib/common/src/dcdCellReader.cpp:26-70: the six unit-cell doubles of
 * each requested frame into outArr float64 [nbrFrames][6]. */
int nda_dcd_get_cell(const char *fileName, const int *frames, int nbrFrames, const long long *startPos,
                     double *outArr, char byteorder);

/* ---- device-pointer entry points (inputs resident in HBM, stream-ordered) ---------- */

/* d_counts uint64 [nGroups][outSize] is ACCUMULATED into (zero it first).
 * d_sel2 == d_sel1 is required when sameSel != 0. */
int nda_dev_radial_nbr_counts(const float *d_sel1, int sel1_size, const float *d_sel2, int sel2_size,
                              const int *d_groupId, int nGroups,
                              unsigned long long *d_counts, int outSize,
                              const float *d_cellDims, int nbrFrames, int frameBegin, int frameEnd,
                              int sameSel, float dr, void *stream);

/* d_bits uint32 [frameEnd-frameBegin][ceil(outSelSize/32)] is overwritten.  If d_out
 * (int32 [nbrAtoms][nbrFrames]) is not NULL, the rows of outSel atoms, columns
 * [frameBegin, frameEnd), are overwritten with 0/1. */
int nda_dev_within(const float *d_allAtoms, int nbrAtoms, int nbrFrames,
                   const int *d_refSel, int refSize, const int *d_outSel, int outSelSize,
                   uint32_t *d_bits, int *d_out, const float *d_cellDims, float distance,
                   int frameBegin, int frameEnd, void *stream);

/* d_sum float64 [sel1_size][sel2_size] is ACCUMULATED into (zero it first); with sameSel
 * only j > i is accumulated (mirror on the host or with nda_dev_symmetrize). */
int nda_dev_distances_sum(const float *d_sel1, int sel1_size, const float *d_sel2, int sel2_size,
                          double *d_sum, const float *d_cellDims, int nbrFrames,
                          int frameBegin, int frameEnd, int sameSel, void *stream);

/* The same with the selections given as row indices into ONE resident all-atom array
 * d_allAtoms float32 [nbrAtoms][nbrFrames][3]: upload a trajectory once, query it many times
 * (B200: 180 GB of HBM holds whole trajectories).  d_groupId has one entry per idx2 entry.
 * sameSel != 0 requires d_idx2 == d_idx1. */
int nda_dev_radial_nbr_counts_sel(const float *d_allAtoms, int nbrAtoms,
                                  const int *d_idx1, int sel1_size, const int *d_idx2, int sel2_size,
                                  const int *d_groupId, int nGroups,
                                  unsigned long long *d_counts, int outSize,
                                  const float *d_cellDims, int nbrFrames, int frameBegin, int frameEnd,
                                  int sameSel, float dr, void *stream);
int nda_dev_distances_sum_sel(const float *d_allAtoms, int nbrAtoms,
                              const int *d_idx1, int sel1_size, const int *d_idx2, int sel2_size,
                              double *d_sum, const float *d_cellDims, int nbrFrames,
                              int frameBegin, int frameEnd, int sameSel, void *stream);

/* ---- measurement / test support ---------------------------------------------------- */

/* Config-5 generator (namdanalyzer_b200/synth.py water_box_frame, bit-identical): writes
 * frames [frameBegin, frameEnd) of an n_side^3-atom box into d_xyz float32
 * [n_side^3][nbrFrames][3] and d_cellDims float32 [nbrFrames][3]. */
int nda_dev_synth_water_box(float *d_xyz, float *d_cellDims, int n_side, int nbrFrames,
                            int frameBegin, int frameEnd, float L, float sigma, uint32_t seed,
                            void *stream);

/* The same generator for a rank that owns frames [firstFrame, firstFrame + nFrames) of a long trajectory:
 * they are written into columns [0, nFrames) of a LOCAL array d_xyz float32 [n_side^3][strideFrames][3]
 * (and rows [0, nFrames) of d_cellDims), bit-identical to the corresponding frames of the full array. */
int nda_dev_synth_water_box_slice(float *d_xyz, float *d_cellDims, int n_side, int strideFrames,
                                  int firstFrame, int nFrames, float L, float sigma, uint32_t seed,
                                  void *stream);

/* Dependent-free FFMA loop on every SM: returns the measured FP32 rate in TFLOP/s
 * (FMA = 2 flop).  variant 0 = scalar FFMA, 1 = packed fma.rn.f32x2. */
int nda_measure_fp32_peak(int variant, double *tflops);

/* Statistics of the last RDF / within launch on the calling thread's context:
 * [0] pairs evaluated by the fast filter, [1] candidates re-evaluated exactly,
 * [2] frames that ran in all-exact mode (cut-off >= half the box), [3] kernel ms. */
int nda_last_stats(double stats[4]);

/* Page-lock (cudaHostRegister) a caller's array so that the host entry points DMA it directly with
 * strided 2-D copies -- no gather through pinned staging, which is bounded by the host's memory
 * bandwidth and does not scale with the number of devices.  The array must stay alive until
 * nda_host_unregister(ptr).  Python: namdanalyzer_b200.pin(array) / unpin(array). */
int nda_host_register(const void *ptr, size_t bytes);
int nda_host_unregister(const void *ptr);

/* Platform ceiling for the end-to-end numbers: `bytes` of contiguous pinned host memory copied to EVERY
 * device of the calling thread at the same time (one plain cudaMemcpyAsync each, `reps` times, one host
 * thread per device); *gbps = aggregate rate over the slowest device's time. */
int nda_measure_h2d(size_t bytes, int reps, double *gbps);

/* Device time (CUDA events on the launching stream) spent in the pair kernel(s) -- rdf_kernel,
 * within_kernel or distances_kernel, one launch per pass -- of the last entry-point call on this
 * device; synchronises on the last of them. */
int nda_last_kernel_ms(double *ms, int *launches);

#ifdef __cplusplus
}
#endif
#endif /* NDA_B200_H */
