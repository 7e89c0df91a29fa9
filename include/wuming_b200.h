/* wuming_b200.h -- C ABI of the B200-native backend for WumingPIC's per-timestep PIC loop.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Each entry point
 * names the reference procedure it replaces (paths relative to the WumingPIC tree).  A Fortran
 * driver binds these through ISO_C_BINDING (see INTEGRATION.md and fortran/wuming_b200_shim.f90);
 * the Python package wumingpic_b200 binds the same symbols through ctypes.
 *
 * Conventions
 *   - every function returns 0 on success and a WM_ERR_* code otherwise; wm_last_error() gives the
 *     message.  The two run-time aborts of the reference are reproduced as error codes:
 *     WM_ERR_CG_ITEMAX   (3d/common/field.f90:522-525, "stop at cgm after ite_max") and
 *     WM_ERR_MEMORY_OVER (3d/common/boundary_periodic.f90:435-438, "memory over (np2 > np)").
 *   - host arrays have exactly the reference's shapes and index bases (column-major):
 *       3-D: up/gp(ndim=7, np, nys:nye, nzs:nze, nsp)   uf(6, nxgs-2:nxge+2, nys-2:nye+2, nzs-2:nze+2)
 *            np2(nys:nye, nzs:nze, nsp)                 cumcnt(nxgs:nxge+1, nys:nye, nzs:nze, nsp)
 *       2-D: up/gp(ndim=6, np, nys:nye, nsp)            uf(6, nxgs-2:nxge+2, nys-2:nye+2)
 *            np2(nys:nye, nsp)                          cumcnt(nxgs:nxge+1, nys:nye, nsp)
 *     (3d/proj/weibel/app.f90:73-77, 282-287; 2d/proj/weibel/app.f90)
 *   - all reals are IEEE binary64, all integers 32-bit, as in the reference (real(8), default integer).
 *   - one wm_ctx per MPI rank / process <-> one GPU.  Calls are stream-ordered and asynchronous
 *     until a download / stats call.  The device copy is authoritative between wm_upload and
 *     wm_download; host arrays are a cache.
 *   - there is NO CPU fallback: every entry point fails with WM_ERR_CUDA if no device is usable.
 */
#ifndef WUMING_B200_H
#define WUMING_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wm_ctx wm_ctx;

enum {
  WM_OK = 0,
  WM_ERR_ARG = 1,          /* bad argument / not initialised ("Initialize first by calling ...__init()") */
  WM_ERR_CUDA = 2,         /* CUDA runtime / NCCL failure, or no device */
  WM_ERR_CG_ITEMAX = 3,    /* cgm did not converge in ite_max = 100 iterations            */
  WM_ERR_MEMORY_OVER = 4,  /* a pencil holds more than np particles (np2 > np)            */
  WM_ERR_PARTICLE_LOST = 5,/* a particle left the one-cell neighbourhood the scheme assumes */
  WM_ERR_STATE = 6         /* call order not supported (e.g. gp download between migrate and sort) */
};

enum { WM_BC_PERIODIC = 0, WM_BC_RECONNECTION = 1, WM_BC_SHOCK = 2 };

/* order of the x-boundary call relative to the field step */
enum { WM_ORDER_WEIBEL = 0,        /* solv, fdtd_i, particle_x, particle_yz, sort   (3d/proj/weibel/app.f90:100-108) */
       WM_ORDER_RECONNECTION = 1,  /* solv, particle_x, fdtd_i, particle_yz, sort   (3d/proj/reconnection/app.f90:103-108) */
       WM_ORDER_SHOCK = 2 };       /* solv, injection, fdtd_i, particle_yz, sort    (2d/proj/shock/app.f90:112-125) */
/* the two pushers of the reference's particle module (3d/common/particle.f90:7) */
enum { WM_PUSHER_BORIS = 0,        /* particle__solv      Buneman-Boris  3d/common/particle.f90:52-233  [2d :48-179]  */
       WM_PUSHER_VAY = 1 };        /* particle__solv_vay  Vay (2008)     3d/common/particle.f90:236-419 [2d :182-315] */

/* Geometry and constants: the union of the arguments of particle__init (3d/common/particle.f90:18-49),
 * field__init (3d/common/field.f90:22-67), sort__init (3d/common/sort.f90:16-37),
 * boundary_periodic__init (3d/common/boundary_periodic.f90:25-65) and the slab ranges / rank grid
 * of mpi_set__init (3d/common/mpi_set.f90:21-97). */
typedef struct wm_params {
  int dim;                 /* 2 or 3 */
  int ndim;                /* 6 (2-D) or 7 (3-D) */
  int np;                  /* pencil capacity of the host arrays */
  int nsp;                 /* number of species (2) */
  int nxgs, nxge, nygs, nyge, nzgs, nzge;  /* global cell index ranges (nz* ignored for dim = 2) */
  int nys, nye, nzs, nze;  /* this rank's slab */
  int nproc_j, nproc_k;    /* rank grid (3-D: rank = j*nproc_k + k; 2-D: nproc_k = 1) */
  int rank_j, rank_k;      /* this rank's coordinates */
  int bc_kind;             /* WM_BC_* */
  int device;              /* CUDA device ordinal, or -1 for the current device */
  double delx, delt, c, gfac;
  double q[2], r[2];       /* charge and mass per species */
} wm_params;

typedef struct wm_stats {
  int cg_iterations[3];    /* iterations of the last cgm call, per component (field.f90 `ite`) */
  long long n_particles;   /* active particles on this rank */
  int max_np2;             /* fullest pencil */
  int error_flags;         /* sticky device-side flags: bit0 memory over, bit1 particle lost, bit2 cg ite_max */
  int timed_steps;         /* steps accumulated in the ms_* sums since wm_set_timing(ctx, 1) */
  double ms_push, ms_deposit, ms_field, ms_sort;  /* CUDA-event time of the phases, SUMMED over timed_steps
                                                     (fused path: ms_push = fused push+deposit kernel, ms_deposit = 0) */
} wm_stats;

const char* wm_last_error(void);
int wm_version(void);

/* -- life cycle --------------------------------------------------------------------------------- */
/* fills nys,nye (and nzs,nze) from the global ranges with para_range (3d/common/mpi_set.f90:81-94) */
int wm_para_range(int n1, int n2, int isize, int irank, int* ns, int* ne);
/* replaces the four __init calls of 3d/proj/weibel/app.f90:341-353 */
int wm_create(const wm_params* prm, wm_ctx** out);
int wm_destroy(wm_ctx* ctx);

/* Multi-GPU: join an NCCL communicator (one rank per GPU).  id_bytes is the 128-byte ncclUniqueId
 * produced by wm_comm_unique_id on rank 0 and broadcast by the host (MPI_Bcast in the Fortran driver,
 * torch.distributed in bench.py).  Replaces the neighbour table of mpi_set__init (mpi_set.f90:63-76). */
int wm_comm_unique_id(char* id_bytes128);
int wm_comm_init(wm_ctx* ctx, int nranks, int rank, const char* id_bytes128);

/* -- state transfer (the sync points of SURVEY.md 8b) ------------------------------------------- */
/* Any pointer may be NULL to skip that array.  up+np2+cumcnt travel together. */
int wm_upload(wm_ctx* ctx, const double* up, const int* np2, const int* cumcnt, const double* uf);
/* up/np2/cumcnt: the sorted particle state; gp: the pushed (unsorted) state, valid between
 * wm_particle_solv and wm_bc_particle_yz; uf: the field incl. ghosts. */
int wm_download(wm_ctx* ctx, double* up, int* np2, int* cumcnt, double* uf, double* gp);
/* internal work arrays for stage-wise parity checks: which = 0 uj(3,box) 1 df(6,box) 2 gkl(3,interior) */
int wm_download_work(wm_ctx* ctx, int which, double* out);
/* inject a work array (which = 1: df, the SAVEd CG warm start of field__fdtd_i, 3d/common/field.f90:102) so that a
 * test can continue from a state the reference/oracle reached; a fresh context starts from df = 0 like the reference */
int wm_upload_work(wm_ctx* ctx, int which, const double* in);

/* -- the hot path on device-resident state, one entry per reference procedure -------------------- */
int wm_particle_solv(wm_ctx* ctx, int nxs, int nxe);    /* particle__solv       3d/common/particle.f90:52-233 [2d :48-179] */
int wm_particle_solv_vay(wm_ctx* ctx, int nxs, int nxe); /* particle__solv_vay  3d/common/particle.f90:236-419 [2d :182-315] */
int wm_field_fdtd_i(wm_ctx* ctx, int nxs, int nxe);     /* field__fdtd_i        3d/common/field.f90:70-208    [2d :66-186]
                                                            incl. ele_cur :211-406, cgm :409-560 and the three boundary
                                                            callbacks (bc__curre, bc__phi, bc__dfield) of ctx's bc_kind  */
int wm_field_stage(wm_ctx* ctx, int nxs, int nxe, int stage); /* one stage of fdtd_i (1 ele_cur 2 curre 3 gkl 4 cgm
                                                            5 dfield 6 dE 7 dfield 8 uf+=df), for stage-wise parity tests */
int wm_bc_particle_x(wm_ctx* ctx, int nxs, int nxe);    /* boundary_*__particle_x  3d/common/boundary_periodic.f90:68-101 */
int wm_bc_injection(wm_ctx* ctx, int nxs, int nxe, double u0); /* boundary_shock__injection 2d/proj/shock/boundary_shock.f90:255-297 */
int wm_bc_particle_yz(wm_ctx* ctx);                     /* boundary_*__particle_y[z] 3d/common/boundary_periodic.f90:104-455 */
int wm_sort_bucket(wm_ctx* ctx, int nxs, int nxe);      /* sort__bucket         3d/common/sort.f90:40-88 */

/* nsteps whole time steps with the fused push+deposit kernel (benchmark path); `order` = WM_ORDER_* */
int wm_step(wm_ctx* ctx, int nxs, int nxe, int order, double u0, int nsteps);
/* 1 (default): wm_step uses the fused push+deposit kernel and the deterministic sort where available;
 * 0: wm_step calls the per-procedure kernels, exactly like a driver calling the five entry points above */
int wm_set_fused(wm_ctx* ctx, int on);
/* wm_step leaves the result of its last sort as (pushed set, permutation): the next wm_step's fused kernel reads through the
 * permutation, which saves one full read + write of the particle store per step.  Every other entry point that reads the sorted
 * set applies the pending permutation first; wm_settle does so explicitly (asynchronous, on the library stream). */
int wm_settle(wm_ctx* ctx);
/* which pusher wm_step / wm_h_step run (WM_PUSHER_*): a driver that calls particle__solv_vay in its time loop sets
 * WM_PUSHER_VAY once; wm_particle_solv and wm_particle_solv_vay always run their own pusher */
int wm_set_pusher(wm_ctx* ctx, int pusher);

/* -- host-buffer forms with the reference's own argument lists (upload, run, download) ---------- */
int wm_h_particle_solv(wm_ctx* ctx, double* gp, const double* up, const double* uf, const int* cumcnt,
                       const int* np2, int nxs, int nxe);
int wm_h_particle_solv_vay(wm_ctx* ctx, double* gp, const double* up, const double* uf, const int* cumcnt,
                           const int* np2, int nxs, int nxe);
int wm_h_field_fdtd_i(wm_ctx* ctx, double* uf, const double* up, const double* gp, const int* cumcnt,
                      const int* np2, int nxs, int nxe);
/* one whole step on host state: up, uf, np2, cumcnt in -> out (gp is scratch on the device only) */
int wm_h_step(wm_ctx* ctx, double* up, double* uf, int* np2, int* cumcnt, int nxs, int nxe, int order, double u0);

/* -- the particle source of the shock set-up on the device --------------------- */
/* Constants of 2d/proj/shock/app.f90's inject() / relocate() / vprofile() (:615-692, 697-852, 883-893; 3d/proj/shock/app.f90
 * :644-728, 733-906, 939-949): n0 particles per cell, upstream flow v0 (< 0), thermal spreads, upstream field, damping length;
 * `seed` keys the Philox streams that replace Fortran's random_number. */
typedef struct wm_shock_params {
  int n0;
  double v0, v_thi, v_the, b0, theta_bn, phi_bn, l_damp_ini;
  unsigned long long seed;
} wm_shock_params;
/* inject(): nlinj[row] new particles per species behind every local pencil, row = (j - nys) + nyl * (k - nzs), placed at
 * x = nxe*delx + (ii - 1/2)/nlinj*|v0|*delt + (v0 + ux)*delt with a drifting Maxwellian (app.f90:786-841); np2 += nlinj,
 * cumcnt(nxe) += nlinj, and the upstream field columns nxe-1, nxe are reset (:840-849).  The integer bookkeeping of
 * :711-781 stays with the caller: nlinj = nlinj_grid, id_first[isp*nrows + row] = ncinj_grid(row) + nptotal(isp), so that
 * particle ii of the row gets ID -(ii + id_first).  `epoch` = the driver's step counter `it`.  Call after sort__bucket. */
int wm_shock_inject(wm_ctx* ctx, const wm_shock_params* prm, int nxe, const int* nlinj, const long long* id_first,
                    long long epoch);
/* relocate() after the caller's nxe = nxe + 1: n0 particles per row and species fill the new cell nxe_new - 1
 * (app.f90:637-676), cumcnt(nxe_new) = cumcnt(nxe_new - 1) + n0; id_first[isp*nrows + row] = global_row * n0 + nptotal(isp) */
int wm_shock_relocate(wm_ctx* ctx, const wm_shock_params* prm, int nxe_new, const long long* id_first, long long epoch);

/* -- particle output on the device -------------------------------------------- */
/* get_particle_count of paraio (3d/common/paraio.f90:1007-1085; the packing behind io__ptcl / io__orb, :843-866): mode 0 packs
 * every active particle, mode 1 the tracers (64-bit ID > 0), species-major in (k, j, cell) order, as consecutive ndim-double
 * records into the HOST buffer buf (capacity cap_records records); lcount[isp] = records of species isp (the routine's lcount).
 * buf may be NULL to obtain the counts only.  Only the packed records cross PCIe. */
int wm_pack_particles(wm_ctx* ctx, int mode, double* buf, long long cap_records, long long* lcount);

/* -- synthetic load and diagnostics on the device ----------------------------------------------- */
/* Weibel load of 3d/proj/weibel/app.f90:311-338,391-504 (2d/proj/weibel/app.f90:404-432) generated on
 * the device with the Philox stream the oracle uses (positions bit-identical, Maxwellian to libm ulp). */
int wm_load_weibel(wm_ctx* ctx, int n0, double v_thi, double v_the, double t_ani, double b0, unsigned long long seed);
/* mom_calc__accl + mom_calc__nvt (3d/common/mom_calc.f90:49-216, 219-332 [2d :49-163, 166-252]) + boundary_*__mom
 * (3d/common/boundary_periodic.f90:1102-1235 [2d :571-636]) on the sorted device-resident particles: the moment block of
 * the drivers (3d/proj/weibel/app.f90:121-124) without downloading a single particle.
 * mom(7, nxgs-1:nxge+1, nys-1:nye+1, [nzs-1:nze+1,] nsp): N, Vx, Vy, Vz, Txx, Tyy, Tzz sums at (i+1/2, j+1/2, k+1/2). */
int wm_mom_calc(wm_ctx* ctx, int nxs, int nxe, double* mom);
/* out[0..nsp-1] kinetic energy per species, out[nsp] E^2/8pi, out[nsp+1] B^2/8pi  (energy_history, app.f90:509-577) */
int wm_energy(wm_ctx* ctx, double* out);
/* out[0] = max|div E - 4 pi rho| , out[1] = max|4 pi rho| over this rank's interior cells */
int wm_gauss(wm_ctx* ctx, double* out);
int wm_get_stats(wm_ctx* ctx, wm_stats* out);
int wm_sync(wm_ctx* ctx);
/* CUDA-event timing of phases inside wm_step: 0 off, 1 on (resets the sums) */
int wm_set_timing(wm_ctx* ctx, int on);
/* the cudaStream_t every kernel of this context is launched on (so that callers can record their own events on it) */
void* wm_stream(wm_ctx* ctx);
/* number of kernels this library launched since wm_create (bench.py's gpu_launches) */
long long wm_launch_count(wm_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* WUMING_B200_H */
