/* w2b_cthyb.h -- C ABI of the B200 CT-HYB engine (libw2b_cthyb.so).
 *
 * Drop-in boundary for the Z-space hot path of w2dynamics' f2py module
 * `w2dyn.auxiliaries.CTQMC.mctqmc` (generated from src/ctqmc_fortran/CTQMC.F90,
 * src/ctqmc_fortran/CMakeLists.txt:54-80).  Every entry point names the reference
 * subroutine it replaces.  Conventions follow what f2py passes to the Fortran code:
 * reals are C double, default integers are int, int64 is long long, arrays are
 * column-major (Fortran order) with the extents of the cited declaration, indices
 * inside arrays (orbitals, spins, ca) are 1-based exactly as in the reference.
 *
 * The reference is a process-wide singleton (module variables, cthyb_solver.py:96-98);
 * so is this library: one solver instance per process, one process per GPU.
 * Unlike the reference (which `stop`s), every call returns 0 on success and a non-zero
 * status on failure; w2b_last_error() returns the message.  There is no CPU fallback:
 * without a usable CUDA device w2b_init_solver fails.
 */
#ifndef W2B_CTHYB_H
#define W2B_CTHYB_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- solver life cycle ------------------------------------------------------------ */
/* init_rnd(nseed)                      CTQMC.F90:9404  (seed of this rank; chains use key (seed, global chain id)) */
int w2b_init_rnd(int nseed);
/* init_paras(parastr)                  CTQMC.F90:9412, grammar Parameters.F90:281-320 */
int w2b_init_paras(const char* parastr, size_t len);
/* init_solver(u_matrix_in, ftau_full, muimp_full, screening_function, iter_no, nftau, nbands)
 *                                      CTQMC.F90:7894-8085
 * u_matrix (2Nb)^4, ftau (Nb,2,Nb,2,Nftau), muimp (Nb,2,Nb,2), screening (Nb,2,Nb,2,Nftau; may be NULL) */
int w2b_init_solver(const double* u_matrix, const double* ftau_full, const double* muimp_full,
                    const double* screening, int iter_no, int nftau, int nbands);
/* set_taudiffmax(taudiffmax)           CTQMC.F90:8180 */
int w2b_set_taudiffmax(double taudiffmax);
/* init_counters()                      CTQMC.F90:847 */
int w2b_init_counters(void);
/* ctqmc_warmup(nwarmups)               CTQMC.F90:8188 */
int w2b_ctqmc_warmup(long long nwarmups);
/* ctqmc_measure(isector, icomponent)   CTQMC.F90:8355 ; only the Z sector (1,1) is implemented */
int w2b_ctqmc_measure(int isector, int icomponent);
/* ctqmc_calibrate(list, phase1pairnum, phase2taugrid, phase2pairnum, nprogress)   CTQMC.F90:8088-8178 ;
 * outputs through w2b_get_array: "accpairtau" (phase 1) and "accpair" (Nb,2,phase2taugrid; phase 2) */
int w2b_ctqmc_calibrate(int list, int phase1pairnum, int phase2taugrid, int phase2pairnum, int nprogress);
/* dest_ctqmc(), dest_paras()           CTQMC.F90:901, 9420 */
int w2b_dest_ctqmc(void);
int w2b_dest_paras(void);

/* ---- Monte-Carlo configurations (CTQMC.F90:9358-9400; per chain, the reference's single chain = chain 0) */
/* outer_sst: 0-based superstate; outer_state: 1-based eigenstate inside it (Trace.F90:6854, 6918) -- the sampled state with
   statesampling = 1, otherwise always 1 (the whole superstate is summed). */
int w2b_get_mc_config_scalars(int chain, int* noper, int* outer_sst, int* outer_state);
int w2b_get_mc_config(int chain, int n, double* taus, int* orbs, int* spins, int* cas, int* hashybs,
                      int* outer_sst, int* outer_state);
int w2b_set_mc_config(int chain, int n, const double* taus, const int* orbs, const int* spins, const int* cas,
                      const int* hashybs, int outer_sst, int outer_state);
int w2b_set_empty_mc_config(int chain, int outer_sst, int outer_state);

/* ---- results: the module variables Python reads after ctqmc_measure (cthyb_solver.py:946-1030) ---- */
/* Writes up to `cap` doubles of the named quantity (column-major, extents as in the reference) and stores the
 * total element count in *n.  Names: gtau gtau_full gleg gleg_full occ histo rhisto lhisto statessuperstates
 * eigenstatesenergies occbasismapping outersuperstatehisto outerstatehisto signhistosuperstates signhistostates
 * tracecontribsuperstates tracecontribstates accadd accrem accpair densitymatrix rho1 rho2 ;
 * gtau_blocks (Nb,2,Ntau,levels) when AccumGtau = 2 (autocorracc, Accumulators.F90:214-364; CTQMC.F90:8733-8739) ;
 * giw (MeasGiw != 0: Giw(NBands, 2, -Niw:Niw-1), CTQMC.F90:573, 8815-8817) as (re, im) pairs = the storage of a complex(8) array ;
 * additionally per-bin samples gtau_bins gleg_bins occ_bins sign_bins giw_bins (leading extents as the quantity, last = bin). */
int w2b_get_array(const char* name, double* out, size_t cap, size_t* n);
/* scalars: mean_sign accglob accshift accadd4 accrem4 accflavc time_warmup time_sim time_sim_steps
 * aborted (0, or after SIGTERM / SIGINT / SIGUSR1 during sampling: 1 = stopped with some measurements, 2 = with none; the traps are armed by
 *          init_solver and restored by dest_ctqmc -- register_trap CTQMC.F90:8076-8078, polls 8214 / 8430, truncation 8648-8658)
 * and engine diagnostics nchains nbins nsst nstates nstatesmax kcap max_k max_ops moves_per_chain gpu_launches cnt_<counter> */
int w2b_get_scalar(const char* name, double* out);

/* ---- rank combination: replaces the per-quantity mpi4py Allreduce of DistributedSample (w2dyn/auxiliaries/statistics.py:95-132)
 * and the one-chain-per-MPI-rank layout (cthyb:144, cthyb_solver.py:319-320).  One process per GPU; chains shard over ranks by
 * global chain id; ctqmc_measure ends with ONE ncclAllReduce(sum, f64) of the raw accumulator slab over NVLink / NVSwitch (plus a
 * 2-int max of error / abort flags), after which every rank normalises and holds the population result.
 * Protocol: rank 0 calls w2b_comm_unique_id (128 bytes), the HOST broadcasts those bytes by its own means (MPI_Bcast from a Fortran
 * or mpi4py host; torch.distributed in bench.py), every rank calls w2b_comm_init(id, 128, rank, nranks) before w2b_init_solver. */
int w2b_comm_unique_id(void* out, size_t cap, size_t* n);
int w2b_comm_init(const void* unique_id, size_t nbytes, int rank, int nranks);
int w2b_comm_destroy(void);
/* the all-reduce by itself (w2b_ctqmc_measure calls it; exposed for the split form w2b_sample / w2b_finalize below) */
int w2b_allreduce_accum(void);

/* ---- engine controls without a reference counterpart ---- */
/* keys: nchains nbins kcap ops_cap chain_base device nmeas_override propose_run scheduler (1x1 blocks: 0/1 = persistent CTA
 * scheduler kernel, 2 = round kernels, 3 = autonomous kernel: one warp takes a chain through all its moves) reuse_population (1: dest_ctqmc parks the chain population and the static tables; the next
 * init_solver of a bit-identical local problem revives them, uploads the new F(tau) and rebuilds every chain's inverse matrices /
 * trace / signs on the device -- ReuseMCConfig (cthyb_solver.py:328-345) for the whole population) ; must be set before w2b_init_solver.  time_kernels (0/1: CUDA events around the round
 * kernels, read back through w2b_get_scalar ms_propose / ms_trace / ms_decide) may be switched at any time. */
int w2b_set_option(const char* key, double value);
/* split form of ctqmc_measure for multi-GPU runs: sample, all-reduce the raw accumulators yourself
 * (device pointer, count in doubles; one NCCL allreduce(sum) over all ranks), then finalize. */
int w2b_sample(long long nmeas, int reset_accumulators);
int w2b_accum_device(void** dev_ptr, size_t* n_doubles);
int w2b_finalize(void);
/* the signal handler itself: a host that owns the process's signal handling forwards SIGTERM / SIGINT / SIGUSR1 here */
void w2b_trap_notify(int sig);
/* one accumulation block: `nsweep` moves per chain followed by one measurement (bench / profiling hook) */
int w2b_block(long long nsweep, int do_measure);
int w2b_sync(void);
/* bench hook: nblocks x (nsweep moves per chain [+ measurement]) timed with CUDA events on the launching stream;
 * ms[0] = whole region, ms[1] = sum over sweep kernels, ms[2] = sum over measurement kernels */
int w2b_bench_blocks(long long nblocks, long long nsweep, int do_measure, double* ms);
/* end-to-end hooks: replace the hybridisation function from HOST memory (layout of init_solver's ftau_full,
 * CTQMC.F90:7950-7955) and copy the raw accumulator slab (nbins x stride doubles) to HOST memory */
int w2b_upload_ftau(const double* ftau_full, int nftau);
int w2b_read_accum(double* out, size_t cap, size_t* n);
/* page-locked host buffers for the two calls above */
void* w2b_pinned_alloc(size_t bytes);
void w2b_pinned_free(void* p);

/* ---- replay / parity hooks ---- */
/* advance all chains by nsteps moves and record chain `chain`'s per-move records (20 doubles each):
 * [0] move 1=add 2=rem 3=glob 4=shift 5=outer add 6=outer rem 7=add4 8=rem4 ([2],[3]: earliest / latest of the four times,
 *     [4] = flavour(pair 1) * nf + flavour(pair 2))  [1] status 0=rejected 1=accepted 2=early exit  [2],[3] tau1,tau2
 * [4] flavour  [5],[6] FPos(c^dag),FPos(c)  [7],[8] log|det ratio|, sign  [9],[10] log|trace ratio|, sign
 * [11] acceptance probability  [12] #operators after  [13] outer superstate after  [14] NPairs used
 * [15] configuration sign after  [16] window position after  [17] accept random number */
int w2b_debug_steps(int chain, long long nsteps, double* log);
/* chain state for parity checks: out[0..7] = trace log, trace sign, det log, det sign, cfgsign, n_ops, outer_sst, error flags */
int w2b_debug_chain_scalars(int chain, double* out);
int w2b_debug_minv(int chain, int orbital, int spin, double* out, size_t cap, int* k);

const char* w2b_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
