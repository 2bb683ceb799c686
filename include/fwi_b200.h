/*
 * fwi_b200.h -- C ABI of the B200-native FWI hot path (libfwi_b200.so).
 *
 * Drop-in boundary of the reference's 2-D elastic FWI operator.  Every entry
 * point below replaces one interface of lidongzh/FwiFlow.jl (paths relative to
 * the reference root):
 *
 *   fwi_b200_cufd      <- cufd(...)       deps/CustomOps/FWI/Src/libCUFD.cu:34-38
 *                                         (declared deps/CustomOps/FWI/FwiOp.h:9-12)
 *   fwi_b200_forward   <- forward(...)    deps/CustomOps/FWI/FwiOp.h:14-19   (calc_id 0)
 *   fwi_b200_backward  <- backward(...)   deps/CustomOps/FWI/FwiOp.h:21-27   (calc_id 1)
 *   fwi_b200_obscalc   <- obscalc(...)    deps/CustomOps/FWI/FwiOp.h:29-33   (calc_id 2)
 *
 * Same argument order, meaning, units and array layouts as the reference:
 *   Lambda, Mu, Den : (nz, nx) double, ROW-MAJOR [z][x] (nz, nx = padded sizes of the
 *                     para file), Lambda/Mu in MPa          (libCUFD.cu:68-78)
 *   stf             : (nShotsTotal, nSteps) double row-major, row = GLOBAL shot id
 *                                                            (Src_Rec.cu:132-137)
 *   shot_ids        : group_size int32, 0-based; survey keys "shot<id>" (Src_Rec.cu:86)
 *   grad_*          : same shapes as the inputs, d(misfit)/d(MPa) for Lambda/Mu;
 *                     grad_stf is (group_size, nSteps), row = POSITION in the group
 *                     (libCUFD.cu:454-456)
 *   para_fname      : path of the single-line JSON parameter file (Parameter.cpp:16-178);
 *                     it names the survey file and the Data directory holding
 *                     Shot<id>.bin (float32, [rec][time])   (libCUFD.cu:189-192,514-521)
 * The only deliberate differences: `const char*` instead of `std::string`, and
 * errors are RETURNED (0 = ok, <0 = error, text via fwi_b200_last_error()) instead
 * of printf + exit() (utilities.h:25-33).
 *
 * All compute runs in hand-written CUDA kernels for sm_100a.  There is no CPU
 * fallback: without a usable CUDA device every compute entry point fails with
 * FWI_B200_ERR_CUDA.
 */
#ifndef FWI_B200_H_
#define FWI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FWI_B200_OK 0
#define FWI_B200_ERR_ARG (-1)     /* bad argument / calc_id                       */
#define FWI_B200_ERR_IO (-2)      /* para / survey / Shot<id>.bin file problem     */
#define FWI_B200_ERR_JSON (-3)    /* malformed or incomplete JSON                  */
#define FWI_B200_ERR_CFL (-4)     /* Courant number > 1 (utilities.cu:225-240)     */
#define FWI_B200_ERR_CUDA (-5)    /* CUDA runtime error / no device                */
#define FWI_B200_ERR_UNSUPPORTED (-6) /* filter / if_src_update requested (if_win IS supported) */
#define FWI_B200_ERR_GEOM (-7)    /* source / receiver outside the grid, grid too small */
/*   (incl. a source or receiver inside the nPad rows below the bottom layer: those rows are never updated and this
 *    implementation neither stores nor loads them; the reference would run with a dead source / an injection nobody sees) */

/* ---- reference-compatible host-buffer entry points ------------------------- */

int fwi_b200_cufd(double *misfit, double *grad_Lambda, double *grad_Mu, double *grad_Den,
                  double *grad_stf, const double *Lambda, const double *Mu, const double *Den,
                  const double *stf, int calc_id, int gpu_id, int group_size,
                  const int *shot_ids, const char *para_fname);

int fwi_b200_forward(double *misfit, const double *Lambda, const double *Mu, const double *Den,
                     const double *stf, int gpu_id, int group_size, const int *shot_ids,
                     const char *para_fname);

int fwi_b200_backward(double *grad_Lambda, double *grad_Mu, double *grad_Den, double *grad_stf,
                      const double *Lambda, const double *Mu, const double *Den,
                      const double *stf, int gpu_id, int group_size, const int *shot_ids,
                      const char *para_fname);

int fwi_b200_obscalc(double *misfit, const double *Lambda, const double *Mu, const double *Den,
                     const double *stf, int gpu_id, int group_size, const int *shot_ids,
                     const char *para_fname);

/* fwi_b200_cufd with two more arguments.  layout: 0 = Lambda / Mu / Den and the three gradient grids are
 * ROW-major [z][x] like the reference's TensorFlow tensors; 1 = COLUMN-major (nz, nx) arrays, element (z, x) at
 * x * nz + z -- what a Julia Matrix is, and the order the device keeps, so neither side transposes (the reference makes
 * Julia transpose into TensorFlow's layout and cufd transpose back, libCUFD.cu:68-78).  stf and grad_stf are unaffected
 * (row = shot).  with_misfit != 0 with calc_id 1: also write *misfit (like fwi_b200_misfit_and_gradient). */
int fwi_b200_cufd_ex(double *misfit, double *grad_Lambda, double *grad_Mu, double *grad_Den,
                     double *grad_stf, const double *Lambda, const double *Mu, const double *Den,
                     const double *stf, int calc_id, int gpu_id, int group_size, const int *shot_ids,
                     const char *para_fname, int layout, int with_misfit);

/* Fused loss + gradient from ONE forward propagation (the reference propagates twice
 * per L-BFGS evaluation: FwiOp.cpp:100 and :220).  Any output pointer may be NULL. */
int fwi_b200_misfit_and_gradient(double *misfit, double *grad_Lambda, double *grad_Mu,
                                 double *grad_Den, double *grad_stf, const double *Lambda,
                                 const double *Mu, const double *Den, const double *stf,
                                 int gpu_id, int group_size, const int *shot_ids,
                                 const char *para_fname);

/* The same evaluation with the shots of the group sharded over `ngpu` devices of this process:
 * shot k of the group goes to gpu_ids[k % ngpu], every device evaluates its shard concurrently (one host thread and
 * one cached plan per device, nothing synchronises), and the packed per-device results [grad_Lambda | grad_Mu |
 * grad_Den | misfit] are summed ON THE DEVICES with one ncclAllReduce (float32, 3 nz nx + 1 values, communicators from
 * ncclCommInitAll, created once per device list; NCCL is bound with dlopen("libnccl.so.2")) on each plan's stream;
 * the reduced buffer is copied to the host once, from the first device.  Replaces the reference's manual
 * sharding over several fwi_op calls with different gpu_id (test/TestFWI.jl:63-69) -- without its doubly counted
 * boundary shots.  grad_stf rows are in the order of shot_ids.  Any output pointer may be NULL.
 * (Across PROCESSES, one per GPU, use the plan API + an NCCL all-reduce of fwi_b200_plan_result_device(): dist.py.) */
int fwi_b200_gradient_multi(double *misfit, double *grad_Lambda, double *grad_Mu, double *grad_Den,
                            double *grad_stf, const double *Lambda, const double *Mu,
                            const double *Den, const double *stf, int ngpu, const int *gpu_ids,
                            int group_size, const int *shot_ids, const char *para_fname);

/* Time-lapse (flow-coupled) FWI in one call: misfit + gradients of `nsurveys` surveys (baseline + monitors), each with
 * its own parameter file / survey file / Data directory and its own model, as the reference's coupled inversion
 * evaluates them with one fwi_op per survey (docs/codes/src_fwi_coupled/main_two_phase_flow_inversion.jl:50-62,84-93).
 * Survey i runs on gpu_ids[i % ngpu]; devices work concurrently (one host thread each), surveys mapped to the same
 * device run back to back on cached plans that stay resident between calls.  All surveys use the shots `shot_ids`
 * and the source time functions `stf`.  misfit[nsurveys]; grad_*[i] may be NULL (per survey or the whole array). */
int fwi_b200_timelapse(int nsurveys, const char *const *para_fnames, const double *const *Lambda,
                       const double *const *Mu, const double *const *Den, const double *stf, int ngpu,
                       const int *gpu_ids, int group_size, const int *shot_ids, double *misfit,
                       double *const *grad_Lambda, double *const *grad_Mu, double *const *grad_Den);

/* Host-only: what the parameter file says (Parameter.cpp:41-144), so that a binding can check the shapes of the
 * caller's arrays before handing pointers over -- the entry points above carry no sizes, exactly like the reference's.
 * out[8] = { nz, nx, nSteps, nPoints_pml, nPad, if_win, scratch_dir_name present, 0 }. */
int fwi_b200_para_info(const char *para_fname, int *out);

/* Developer A/B switches (not part of the reference's surface).  "rev_lean": -1 pick the build of the reverse-time
 * kernel by working-set size (default), 0 / 1 force the double-buffered / LEAN build.  "merged_bwd": 0 (default) the
 * backward loop is two launches per time index (reverse step / imaging, adjoint step), 1 one merged launch (moves fewer
 * DRAM bytes but measured slower: latency-bound).  "frame_ring": depth of the boundary ring saved per
 * time step for the reverse-time reconstruction, for plans created afterwards: 2 (default) the two cells outside the
 * inner box that the stencils of the box cells read, 5 the reference's ring (Boundary.cu:17-27: those two plus three
 * box cells); same gradients to 2e-5, 2.1x the frame bytes.  "acc_group": shots of a launch that share one imaging
 * accumulator slot in the reverse step (the CTA of a tile takes them one after the other and keeps the sums in shared
 * memory): 0 (default) chosen by working-set size -- groups of <= 12 on the DRAM-bound grids, 1 elsewhere --, k >= 1
 * forced; same gradients up to the order of the float sums (1e-6), deterministic for a given value.  "dyn_units": 1
 * (default) the reverse step's (tile, shot group) units are claimed from a counter in device memory, 0 dealt
 * round-robin (same results; measured slower on long launches). */
int fwi_b200_set_option(const char *name, int value);

/* Host-only: the device layout this library derives from a parameter file (no GPU needed).
 * out[12] = { nz, nx, column pitch, zlive (rows >= zlive are never stored), z_off (row of the first forward/adjoint
 * tile, <= 0, multiple of 4), tiles_z, tiles_x, floats per field of one saved boundary frame, zlo, zhi, xlo, xhi
 * (inner box = reconstruction / imaging region, Boundary.cu:17-27) }. */
int fwi_b200_grid_info(const char *para_fname, int *out);

/* Text of the last error raised on the calling thread ("" if none). */
const char *fwi_b200_last_error(void);

/* Free every cached device context (plans are cached per gpu / para file content). */
void fwi_b200_release(void);

/* ---- device-resident plan API (what the host entry points are built on) ----- *
 * A plan owns, on one GPU, everything that `cufd` allocates and frees per call
 * (libCUFD.cu:115-137,538-575): model + derived coefficients, CPML profiles,
 * wavefields for a batch of concurrent shots, boundary-frame store, traces.
 * Inputs can be pushed once and kept resident across calls; outputs stay on the
 * device until asked for, so that a multi-GPU driver can all-reduce them in place. */
typedef struct fwi_b200_plan fwi_b200_plan;

/* max_batch: shots advanced per kernel launch (0 = choose from free memory). */
int fwi_b200_plan_create(fwi_b200_plan **plan, const char *para_fname, int gpu_id,
                         int group_size, const int *shot_ids, int max_batch);
void fwi_b200_plan_destroy(fwi_b200_plan *plan);

/* host -> device.  Same layouts as fwi_b200_cufd. */
int fwi_b200_plan_set_model(fwi_b200_plan *plan, const double *Lambda, const double *Mu,
                            const double *Den);
int fwi_b200_plan_set_stf(fwi_b200_plan *plan, const double *stf);
/* layout of the (nz, nx) grids this plan takes (set_model) and returns (get_result, result_device): 0 row-major [z][x]
 * (default), 1 column-major.  Changing it invalidates the resident model. */
int fwi_b200_plan_set_layout(fwi_b200_plan *plan, int layout);
/* Velocity-space front end: what FWI.jl does in TensorFlow around the op (src/FWI.jl:156-205, src/Utils.jl:221-227),
 * done on the device.  cp, cs [m/s], rho [kg/m^3] in the plan's layout; padded = 0: (nz - 2 nPml - nPad, nx - 2 nPml)
 * grids, extended like tf.pad(..., "SYMMETRIC") (FWI.jl:166-170); padded = 1: (nz, nx) grids.  is_masked = 0 blends
 * with the padded reference models outside the reference's mask (FWI.jl:45-49, 174-176; *_ref required, same shape
 * as cp), is_masked != 0 takes the models as they are (*_ref ignored).  Then lambda = (cp^2 - 2 cs^2) rho / 1e6,
 * mu = cs^2 rho / 1e6 (Utils.jl:221-227) and the Courant check of set_model. */
int fwi_b200_plan_set_velocities(fwi_b200_plan *plan, const double *cp, const double *cs, const double *rho,
                                 const double *cp_ref, const double *cs_ref, const double *rho_ref,
                                 int is_masked, int padded);
/* After a calc_id 1 run on a model given by set_velocities: the misfit and d misfit / d (cp, cs, rho) on the PADDED
 * (nz, nx) grid in the plan's layout -- the chain rule TensorFlow applies through velocity_to_moduli and the mask
 * blend (zero outside the mask when is_masked was 0).  Gradients w.r.t. unpadded inputs are the fold of these over
 * the symmetric padding (the adjoint of tf.pad), left to the caller. */
int fwi_b200_plan_get_velocity_gradients(fwi_b200_plan *plan, double *misfit, double *g_cp, double *g_cs,
                                         double *g_rho);
/* observed data of the i-th shot of the group: (nrec, nSteps) float32, time fastest. */
int fwi_b200_plan_set_obs(fwi_b200_plan *plan, int ishot, const float *obs);
/* read data_dir_name/Shot<id>.bin for every shot of the group (libCUFD.cu:189-192). */
int fwi_b200_plan_load_obs_files(fwi_b200_plan *plan);

/* Enqueue one evaluation on `stream` (a cudaStream_t; NULL = the plan's own stream):
 * calc_id 0 misfit, 1 gradient (+misfit), 2 synthetic traces.  Does not synchronise
 * unless `sync` != 0.  Results stay on the device. */
int fwi_b200_plan_run(fwi_b200_plan *plan, int calc_id, void *stream, int sync);

/* Device pointer to the packed result [grad_Lambda | grad_Mu | grad_Den | misfit]:
 * 3*nz*nx + 1 float32, gradients ROW-MAJOR [z][x] per MPa, misfit = 0.5*sum(res^2).
 * This is the buffer a multi-GPU driver all-reduces (ncclAllReduce, sum). */
float *fwi_b200_plan_result_device(fwi_b200_plan *plan);
size_t fwi_b200_plan_result_count(fwi_b200_plan *plan);

/* device -> host copies (synchronise the plan's stream first). */
int fwi_b200_plan_get_result(fwi_b200_plan *plan, double *misfit, double *grad_Lambda,
                             double *grad_Mu, double *grad_Den, double *grad_stf);
/* which: 0 synthetic, 1 residual (tapered), 2 conditioned observed -- (nrec, nSteps) float32 */
int fwi_b200_plan_get_traces(fwi_b200_plan *plan, int ishot, int which, float *out);
int fwi_b200_plan_write_obs_files(fwi_b200_plan *plan);

/* Geometry / bookkeeping queries. */
int fwi_b200_plan_info(fwi_b200_plan *plan, int *nz, int *nx, int *nSteps, int *nPml, int *nPad,
                       int *group_size, int *batch, int *max_nrec);
/* resolved 0-based padded indices of the i-th shot of the group (Src_Rec.cu:86-113). */
int fwi_b200_plan_shot_geometry(fwi_b200_plan *plan, int ishot, int *z_src, int *x_src,
                                int *nrec, int *z_rec, int *x_rec);
/* number of kernels launched by the plan since creation (bench.py's gpu_launches). */
long long fwi_b200_plan_launch_count(fwi_b200_plan *plan);
/* Debug / invariant tests: copy one wavefield of shot `ishot` of the current batch to the
 * host as [z][x] float32.  field: 0 vz 1 vx 2 szz 3 sxx 4 sxz (forward / reconstructed). */
int fwi_b200_plan_get_field(fwi_b200_plan *plan, int ishot, int field, float *out);

/* Kernel timing for the roofline: runs `iters` launches of one hot kernel on the plan's
 * current state (which: 0 forward step, 1 forward step + frame save, 2 reverse+imaging,
 * 3 adjoint step, 4 merged backward launch = adjoint step + reverse/imaging) on `stream` between two CUDA events,
 * returns the mean ms per launch
 * and the algorithmic bytes one launch moves. */
int fwi_b200_plan_time_kernel(fwi_b200_plan *plan, int which, int iters, void *stream,
                              float *ms_per_launch, double *alg_bytes_per_launch);

/* Library / device self-description (JSON text, static storage). */
const char *fwi_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FWI_B200_H_ */
