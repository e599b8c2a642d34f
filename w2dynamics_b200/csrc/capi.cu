// C ABI of the engine (see include/w2b_cthyb.h) -- a process-wide singleton like the reference's
// Fortran module (CTQMC.F90 module variables; cthyb_solver.py:96-98).
#include <chrono>
#include <csignal>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <dlfcn.h>
#if defined(__CUDACC__) && !defined(W2B_HOST_EMU)
#include <nccl.h>
#endif

#include "../../include/w2b_cthyb.h"
#include "engine.hpp"
#include "params.hpp"
#include "postproc.hpp"

namespace {

using namespace w2b;

struct Solver {
    Params par;
    bool have_par = false;
    int seed = 0;
    std::unique_ptr<Engine> eng;
    EngineOptions opt;
    long long nmeas_override = -1;
    int device = -1;
    int same_key = 0;
    int prefetch = 1;
    int four_rank2 = 1;
    int time_kernels = 0;
    int propose_run = 4;
    // derived
    double beta = 0; int nb = 0; long long nmeas = 0; int ncorr = 100;
    double cfg_taudiffmax = -1;
    std::vector<int> dens_off;
    ResultSet res; bool have_res = false;
    std::vector<std::vector<double>> bin_gtau, bin_gleg, bin_occ, bin_giw; std::vector<double> bin_sign;
    std::vector<double> raw_acc; bool bins_ready = false;      // raw accumulator slabs of the last finalize; per-bin results are normalised on first request
    double time_warmup = 0, time_sim = 0, time_sim_steps = 0;
    Rho2Cache rho2cache; bool rho2_ready = false;
    unsigned nsolve = 0;            // init_solver calls since init_rnd (selects the section of every chain's RNG stream)
    int aborted = 0;                // 0 / 1 (some measurements done) / 2 (none): CTQMC.F90:8648-8658
    long long nmeas_done = 0;
    // population kept across dest_ctqmc / init_solver (option reuse_population): static tables + thermalised chains
    std::unique_ptr<Engine> parked; std::vector<double> key_u, key_mu; std::string key_par; EngineOptions key_opt; int key_nftau = 0;
    int reuse_population = 0; long long key_nmeas = -1;
    std::vector<double> accpairtau, calib_accpair;     // ctqmc_calibrate outputs
    std::string err;
};
Solver S;

int fail(const std::string& m) { S.err = m; return 1; }

// ---- signal traps (register_trap(SIGTERM | SIGINT | SIGUSR1, trap_notify), CTQMC.F90:8076-8078; Signals.F90:117-137)
volatile std::sig_atomic_t g_signal = 0;
const int TRAP_SIGS[3] = {SIGTERM, SIGINT, SIGUSR1};
struct sigaction g_old_sa[3];
bool g_traps_armed = false;
}  // namespace
extern "C" void w2b_trap_notify(int sig);
namespace {
void arm_traps() {
    if (g_traps_armed) return;
    struct sigaction sa; std::memset(&sa, 0, sizeof sa);
    sa.sa_handler = w2b_trap_notify; sigemptyset(&sa.sa_mask);
    for (int i = 0; i < 3; ++i) sigaction(TRAP_SIGS[i], &sa, &g_old_sa[i]);
    g_traps_armed = true;
}
void disarm_traps() {
    if (!g_traps_armed) return;
    for (int i = 0; i < 3; ++i) sigaction(TRAP_SIGS[i], &g_old_sa[i], nullptr);
    g_traps_armed = false;
}
bool signal_fired() { return g_signal != 0; }

// ---- rank combination inside the library: chains shard over ranks, one all-reduce per ctqmc_measure
struct Comm {
    int rank = 0, nranks = 1;
    void (*hook)(double*, size_t, int*) = nullptr;            // host-emulation build only (tests): stands in for NCCL
#if W2B_DEVICE
    void* lib = nullptr; ncclComm_t comm = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    void load() {
        if (lib) return;
        // the process may already hold an NCCL (torch bundles one): use that instance, otherwise the system library
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) throw std::runtime_error(std::string("cannot load libnccl.so.2: ") + dlerror());
        auto sym = [&](const char* n) { void* p = dlsym(lib, n); if (!p) throw std::runtime_error(std::string("libnccl: missing symbol ") + n); return p; };
        GetUniqueId = (decltype(GetUniqueId))sym("ncclGetUniqueId"); CommInitRank = (decltype(CommInitRank))sym("ncclCommInitRank");
        AllReduce = (decltype(AllReduce))sym("ncclAllReduce"); CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
    }
    void ck(ncclResult_t r, const char* what) { if (r != ncclSuccess) throw std::runtime_error(std::string("NCCL ") + what + ": " + GetErrorString(r)); }
#endif
};
Comm G;
double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template <class F> int guarded(F f) {
    try { f(); return 0; } catch (const std::exception& e) { return fail(e.what()); }
}

void require_engine() { if (!S.eng) throw std::runtime_error("solver not initialised (call w2b_init_solver)"); }

void check_chain(int chain) {
    require_engine();
    if (chain < 0 || chain >= S.eng->C.nchains) throw std::runtime_error("chain index " + std::to_string(chain) + " outside [0, " + std::to_string(S.eng->C.nchains) + ")");
}

void check_supported(const Params& p) {
    auto zero_i = [&](const char* k) { if (p.has(k) && p.integer(k) != 0) throw std::runtime_error(std::string(k) + " != 0 is outside the implemented hot path"); };
    auto zero_r = [&](const char* k) { if (p.has(k) && p.real(k) != 0.0) throw std::runtime_error(std::string(k) + " != 0 is outside the implemented hot path"); };
    zero_i("segment"); zero_i("offdiag"); zero_i("Uw"); zero_i("Phonon");
    zero_i("MeasG2iw"); zero_i("MeasSusz"); zero_i("MeasExpResDensityMatrix"); zero_i("FourPnt");
    zero_i("MeasSsquare"); zero_i("MeasGtauDetRat"); zero_i("flavourchange_moves"); zero_i("Gtau_mean_step"); zero_i("Gtau_mid_step"); zero_i("sign_step");
    // PercentageOuterMove (outer pair moves, CTQMC.F90:3422-3449, 3934-3961) and Percentage4OperatorMove (StepAdd4 / StepRem4 with
    // diagonal hybridisation, CTQMC.F90:3615-3889, 4153-4430) are implemented
    if (p.has("Percentage4OperatorMove") && (p.real("Percentage4OperatorMove") < 0.0 || p.real("Percentage4OperatorMove") > 1.0)) throw std::runtime_error("Percentage4OperatorMove must lie in [0, 1]");
    if (p.has("PercentageOuterMove") && (p.real("PercentageOuterMove") < 0.0 || p.real("PercentageOuterMove") > 1.0)) throw std::runtime_error("PercentageOuterMove must lie in [0, 1]");
    if (p.has("Eigenbasis") && p.integer("Eigenbasis") == 0) throw std::runtime_error("You requested Krylov, but this no longer supported.");
    if (p.has("EnableOuterTruncation") && p.str("EnableOuterTruncation") == "YES") throw std::runtime_error("EnableOuterTruncation is outside the implemented hot path");
    if (p.has("AccumGtau") && p.integer("AccumGtau") != 1 && p.integer("AccumGtau") != 2) throw std::runtime_error("AccumGtau must be 1 (mean) or 2 (mean + autocorrelation blocks)");
}

void finalize_results() {
    require_engine();
    Engine& E = *S.eng;
    std::vector<double> acc; std::vector<unsigned long long> accpair;
    E.download_accum(acc, accpair);
    const int nbins = E.R.nbins; const long long st = E.accum_doubles;
    std::vector<double> tot(st, 0.0);
    for (int b = 0; b < nbins; ++b) for (long long i = 0; i < st; ++i) tot[i] += acc[(size_t)b * st + i];
    if (tot[E.A.o_cnt] <= 0.0) throw std::runtime_error("No measurements");
    // where the reference `stop`s (CTQMC.F90:4062 small determinant ratio; array bounds it never checks) a chain raises a flag;
    // results of a population with a flagged chain are not handed out
    {
        int mk = 0, mo = 0; const int e = E.error_flags(&mk, &mo);
        if (e) {
            std::string m = "chain error flags:";
            if (e & ERR_CAPACITY) m += " operator / matrix capacity exceeded (max order " + std::to_string(mk) + " of kcap " + std::to_string(E.C.kcap) + ", max operators " + std::to_string(mo) + " of ops_cap " + std::to_string(E.C.ops_cap) + "; raise the kcap / ops_cap options);";
            if (e & ERR_DETRAT_SMALL) m += " accepted removal with |determinant ratio| < 1e-12 (CTQMC.F90:4062);";
            if (e & ERR_PAIR_LOGIC) m += " pair enumeration inconsistent;";
            if (e & ERR_GLOBAL_QN) m += " no outer superstate compatible with the configuration;";
            if (e & ERR_SINGULAR) m += " singular hybridisation matrix;";
            throw std::runtime_error(m);
        }
    }
    S.res = ResultSet();
    if (!S.rho2_ready) { S.rho2cache.build(E.at); S.rho2_ready = true; }
    normalise_slab(E.at, E.A, E.R, S.beta, tot.data(), S.dens_off, true, S.res, &S.rho2cache);
    // per-bin samples (for error bars across chain groups) are normalised on first request (w2b_get_array "*_bins")
    S.raw_acc.swap(acc); S.bins_ready = false;
    S.bin_gtau.clear(); S.bin_gleg.clear(); S.bin_occ.clear(); S.bin_giw.clear(); S.bin_sign.clear();
    // move statistics (CTQMC.F90:8680-8695)
    std::vector<unsigned long long> cnt; E.get_counters(cnt);
    auto ratio = [](unsigned long long a, unsigned long long b) { return b ? (double)a / (double)b : std::numeric_limits<double>::quiet_NaN(); };
    std::vector<double> accadd(16, 0.0), accrem(16, 0.0);
    accadd[0] = ratio(cnt[C_ACCADD], cnt[C_TRYADD]); accrem[0] = ratio(cnt[C_ACCREM], cnt[C_TRYREM]);
    S.res.arr["accadd"] = accadd; S.res.arr["accrem"] = accrem;
    S.res.scal["accglob"] = ratio(cnt[C_ACCGLOB], cnt[C_TRYGLOB]); S.res.scal["accshift"] = ratio(cnt[C_ACCSHIFT], cnt[C_TRYSHIFT]);
    S.res.scal["four_rank2_updates"] = (double)cnt[C_FOUR_RANK2];
    S.res.scal["accadd4"] = ratio(cnt[C_ACCADD4], cnt[C_TRYADD4]); S.res.scal["accrem4"] = ratio(cnt[C_ACCREM4], cnt[C_TRYREM4]);   // AccAdd4 / TryAdd4: NaN without 4-operator moves, as in the reference
    S.res.scal["accflavc"] = std::numeric_limits<double>::quiet_NaN();
    S.res.scal["steps"] = (double)cnt[C_STEPS];
    if (E.ac_levels) {                                                       // gtau_blocks(NBands, 2, NGtau, levels), raw as in the reference (CTQMC.F90:8733-8739)
        std::vector<double> blk; E.autocorr_blocks(blk);
        const int nb = E.at.nb, nf = E.at.nf, ntau = E.R.ntau;
        std::vector<double> gb((size_t)nf * ntau * E.ac_levels);
        for (int l = 0; l < E.ac_levels; ++l) for (int b = 0; b < nb; ++b) for (int s2 = 0; s2 < 2; ++s2) for (int t = 0; t < ntau; ++t)
            gb[(b + (size_t)nb * s2) + (size_t)nf * t + (size_t)nf * ntau * l] = blk[(size_t)l * nf * ntau + (size_t)(2 * b + s2) * ntau + t];
        S.res.arr["gtau_blocks"] = gb;
        S.res.scal["gtau_levels"] = (double)E.ac_levels;
    }
    std::vector<double> ap(accpair.size());
    const int nf = E.T.nf, nbd = E.T.nb, ntau = E.R.ntau;
    for (int b = 0; b < nbd; ++b) for (int s = 0; s < 2; ++s) for (int t = 0; t < ntau; ++t) ap[b + (size_t)nbd * (s + 2 * t)] = (double)accpair[(size_t)(2 * b + s) * ntau + t];
    S.res.arr["accpair"] = ap;
    // static tables Python also reads (CTQMC.F90:8053-8066)
    const AtomicTables& at = E.at;
    std::vector<double> sss(at.nstates), ene(at.nstates), occm((size_t)at.nstates * nbd * 2);
    for (int k = 0; k < at.nsst; ++k) for (int i = 0; i < at.dim[k]; ++i) {
        int idx = at.offset[k] + i; sss[idx] = k; ene[idx] = at.evals[idx];
        int st_ = at.block_states[k][i];
        for (int b = 0; b < nbd; ++b) for (int s = 0; s < 2; ++s) occm[idx + (size_t)at.nstates * (b + (size_t)nbd * s)] = (st_ >> (b + s * nbd)) & 1;
    }
    S.res.arr["statessuperstates"] = sss; S.res.arr["eigenstatesenergies"] = ene; S.res.arr["occbasismapping"] = occm;
    S.res.arr["gtau_full"] = std::vector<double>((size_t)nf * nf * ntau, 0.0);
    S.res.arr["gleg_full"] = std::vector<double>((size_t)nf * nf * E.R.nlegmax, 0.0);
    S.res.arr["rhisto"] = std::vector<double>(100, 0.0);
    S.res.arr["lhisto"] = std::vector<double>(100, std::numeric_limits<double>::quiet_NaN());   // 0/0 in the reference (CTQMC.F90:8698)
    S.have_res = true;
}

// the parts of the parameter list that shape the engine (everything but the run lengths)
std::string key_of(const Params& p) {
    std::string k;
    for (auto& kv : p.items) if (kv.first != "Nmeas" && kv.first != "Nwarmups" && kv.first != "Nwarmups2Plus") k += kv.first + "=" + kv.second + ";";
    return k;
}
bool same_opt(const EngineOptions& a, const EngineOptions& b) {
    return a.nchains == b.nchains && a.nbins == b.nbins && a.kcap == b.kcap && a.ops_cap == b.ops_cap && a.cta_warps == b.cta_warps &&
           a.batch_chains == b.batch_chains && a.scheduler == b.scheduler && a.lu_groups == b.lu_groups && a.seed == b.seed && a.chain_base == b.chain_base;
}

void bins_materialise() {
    if (S.bins_ready) return;
    require_engine();
    Engine& E = *S.eng;
    const int nbins = E.R.nbins; const long long st = E.accum_doubles;
    for (int b = 0; b < nbins; ++b) {
        const double* slab = S.raw_acc.data() + (size_t)b * st;
        if (slab[E.A.o_cnt] <= 0.0) continue;
        ResultSet r; normalise_slab(E.at, E.A, E.R, S.beta, slab, S.dens_off, false, r);
        S.bin_gtau.push_back(r.arr["gtau"]); S.bin_gleg.push_back(r.arr["gleg"]); S.bin_occ.push_back(r.arr["occ"]);
        if (r.arr.count("giw")) S.bin_giw.push_back(r.arr["giw"]);
        S.bin_sign.push_back(r.scal["mean_sign"]);
    }
    S.bins_ready = true;
}

// capacity / consistency flags checked while the run is young instead of after the last measurement
void check_flags_early(const char* when) {
    Engine& E = *S.eng;
    int mk = 0, mo = 0; const int e = E.error_flags(&mk, &mo);
    if (e & ERR_CAPACITY) throw std::runtime_error(std::string("operator / matrix capacity exceeded ") + when + " (max order " + std::to_string(mk) + " of kcap " + std::to_string(E.C.kcap) +
                                                   ", max operators " + std::to_string(mo) + " of ops_cap " + std::to_string(E.C.ops_cap) + "): raise the kcap / ops_cap options");
}

void run_blocks(long long nmeas, bool timing) {
    Engine& E = *S.eng;
#if W2B_DEVICE
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float steps_ms = 0.f;
#endif
    S.nmeas_done = 0;
    for (long long i = 0; i < nmeas; ++i) {
        if (signal_fired()) break;                               // CTQMC.F90:8430: `if (any_signal_fired) exit OUTER_LOOP`
        if (i > 0 && (i & 63) == 0) check_flags_early("while sampling");
#if W2B_DEVICE
        if (timing) cudaEventRecord(e0);
#endif
        E.sweep(S.ncorr);
#if W2B_DEVICE
        if (timing) cudaEventRecord(e1);
#endif
        E.measure();
        E.autocorr_add();
#if W2B_DEVICE
        if (timing) { cudaEventSynchronize(e1); float ms = 0; cudaEventElapsedTime(&ms, e0, e1); steps_ms += ms; }
#endif
        S.nmeas_done = i + 1;
    }
    E.sync();
    if (signal_fired()) {                                        // CTQMC.F90:8648-8658: NMeas = iNMeas - 1; aborted = 2 (no data) / 1
        std::fprintf(stderr, "Rank%3d> WARNING: Process caught signal\nRank%3d; measurement%10lld of%10lld\n", G.rank, G.rank, S.nmeas_done + 1, nmeas);
        S.aborted = S.nmeas_done == 0 ? 2 : 1;
    }
#if W2B_DEVICE
    if (timing) S.time_sim_steps += steps_ms * 1e-3;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
#endif
}

}  // namespace

extern "C" {

const char* w2b_last_error(void) { return S.err.c_str(); }
void w2b_trap_notify(int sig) { g_signal = sig ? sig : 1; }

int w2b_init_rnd(int nseed) { S.seed = nseed; S.nsolve = 0; return 0; }

int w2b_init_paras(const char* parastr, size_t len) {
    return guarded([&] { S.par.parse(std::string(parastr, len)); S.have_par = true; check_supported(S.par); });
}

int w2b_set_option(const char* key, double value) {
    std::string k(key);
    if (k == "nchains") S.opt.nchains = (int)value;
    else if (k == "nbins") S.opt.nbins = (int)value;
    else if (k == "kcap") S.opt.kcap = (int)value;
    else if (k == "ops_cap") S.opt.ops_cap = (int)value;
    else if (k == "chain_base") S.opt.chain_base = (unsigned)value;
    else if (k == "cta_warps") S.opt.cta_warps = (int)value;
    else if (k == "batch_chains") S.opt.batch_chains = (int)value;
    else if (k == "scheduler") S.opt.scheduler = (int)value;
    else if (k == "lu_groups") S.opt.lu_groups = (int)value;
    else if (k == "device") S.device = (int)value;
    else if (k == "nmeas_override") S.nmeas_override = (long long)value;
    else if (k == "debug_same_key") S.same_key = (int)value;
    else if (k == "prefetch") S.prefetch = (int)value;
    else if (k == "four_rank2") S.four_rank2 = (int)value;
    else if (k == "time_kernels") { S.time_kernels = (int)value; if (S.eng) { S.eng->time_kernels = value != 0; S.eng->ms_propose = S.eng->ms_trace = S.eng->ms_decide = 0; } }
    else if (k == "propose_run") S.propose_run = (int)value;
    else if (k == "reuse_population") S.reuse_population = (int)value;
    else return fail("unknown option " + k);
    return 0;
}

int w2b_init_solver(const double* u_matrix, const double* ftau_full, const double* muimp_full, const double* screening,
                    int iter_no, int nftau, int nbands) {
    // screening_function enters the reference only with Screening / Uw != 0, which init_paras rejects; iter_no only labels its
    // output files (CTQMC.F90:7894-7948).  Successive solves draw fresh random numbers through the solve counter below.
    (void)screening; (void)iter_no;
    return guarded([&] {
        if (!S.have_par) throw std::runtime_error("init_paras must be called before init_solver");
        const Params& p = S.par;
#if W2B_DEVICE
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw std::runtime_error("no CUDA device available: the B200 engine has no CPU fallback");
        if (S.device >= 0) W2B_CUDA_CHECK(cudaSetDevice(S.device));
#endif
        S.nb = (int)p.integer("Nd");
        if (S.nb != nbands) throw std::runtime_error("Nd in the parameter string does not match the array extents");
        S.beta = p.real("beta");
        S.nmeas = S.nmeas_override >= 0 ? S.nmeas_override : p.integer("Nmeas");
        S.ncorr = (int)p.integer("NCorr");
        S.cfg_taudiffmax = p.real("TaudiffMax");
        const int nb = S.nb, nf = 2 * nb;
        // diagonal hybridisation (CTQMC.F90:7950-7955): Ftau(b,s,:) = Ftau_full(b,s,b,s,:)
        std::vector<double> fd((size_t)nf * nftau);
        for (int b = 0; b < nb; ++b) for (int s = 0; s < 2; ++s) for (int t = 0; t < nftau; ++t)
            fd[(size_t)(2 * b + s) * nftau + t] = ftau_full[b + (size_t)nb * (s + 2 * (b + (size_t)nb * (s + 2 * (size_t)t)))];
        Run R{};
        R.pglob = p.real("PercentageGlobalMove"); R.pshift = p.real("PercentageTauShiftMove");
        R.p4op = p.real_or("Percentage4OperatorMove", 0.0); R.pouter = p.real_or("PercentageOuterMove", 0.0);
        R.statesampling = (p.has("statesampling") && p.integer("statesampling") != 0) ? 1 : 0;
        R.ngiw = (p.has("MeasGiw") && p.integer("MeasGiw") != 0) ? (int)p.integer("Niw") : 0;       // Giw(NBands, 2, -NGiw:NGiw-1), CTQMC.F90:300, 569-575
        R.ntau = (int)p.integer("Ntau"); R.nlegmax = (int)p.integer("NLegMax"); R.maxhisto = (int)p.integer("MaxHisto");
        R.has_szt = p.str("QuantumNumbers").find("Szt") != std::string::npos;
        R.meas_dm = p.integer("MeasDensityMatrix") != 0;
        R.nbins = std::max(1, std::min(S.opt.nbins, S.opt.nchains));
        EngineOptions opt = S.opt;
        if (G.nranks > 1) opt.chain_base += (unsigned)G.rank * (unsigned)opt.nchains;      // chains shard over ranks by global chain id
        R.seed = (unsigned)S.seed; R.chain_base = opt.chain_base; R.same_key = S.same_key; R.prefetch = S.prefetch; R.four_rank2 = S.four_rank2; R.propose_run = S.propose_run > 0 ? S.propose_run : 1;
        R.stream = S.nsolve++;
        // user symmetry moves (symm_to_pstring cthyb_solver.py:769-774; init_Trace Trace.F90:403-404, 471-478): NSymMove lists of 2 Nb
        // target indices (b + Nb s + 1), followed by their inverses
        std::vector<int> symmoves;
        R.nsym = (int)p.integer_or("NSymMove", 0);
        if (R.nsym > 0) {
            symmoves.assign((size_t)2 * R.nsym * nf, 0);
            for (int i = 0; i < R.nsym; ++i) {
                char id[16]; std::snprintf(id, sizeof id, "SymMove%02d", i + 1);
                const std::vector<int> v = p.integer_list(id, nf);
                std::vector<int> seen(nf + 1, 0);
                for (int j = 0; j < nf; ++j) {
                    if (v[j] < 1 || v[j] > nf || seen[v[j]]++) throw std::runtime_error(std::string(id) + " is not a permutation of 1.." + std::to_string(nf));
                    symmoves[(size_t)i * nf + j] = v[j];
                    symmoves[(size_t)(R.nsym + i) * nf + v[j] - 1] = j + 1;
                }
            }
        }
        // init_window with the configured TaudiffMax, NSlide fixed here (CTQMC.F90:8068-8069)
        if (S.cfg_taudiffmax > 0.0) { R.taudiff_max = S.cfg_taudiffmax; R.num_windows = std::max((int)std::floor(S.beta / S.cfg_taudiffmax) - 1, 1); }
        else { R.taudiff_max = S.beta; R.num_windows = 1; }
        R.nslide = std::max(2 * nb, (int)std::floor((double)S.ncorr / (double)R.num_windows));
        opt.seed = (unsigned)S.seed;
        const size_t nu = (size_t)nf * nf * nf * nf, nmu = (size_t)nf * nf;
        // a population parked by dest_ctqmc (option reuse_population) is revived when the local problem is bit-identical: static
        // tables and thermalised chains stay on the device, only F(tau) is replaced and every chain's inverse matrices,
        // determinant, trace and signs are rebuilt from its operator list with the new F (what set_mc_config does for one chain,
        // Trace.F90:6872-7000) -- ReuseMCConfig for the whole population
        const bool blocks = p.has("AccumGtau") && p.integer("AccumGtau") == 2;   // its level count follows Nmeas
        const bool revive = S.reuse_population && S.parked && S.key_par == key_of(p) && S.key_nftau == nftau && same_opt(S.key_opt, opt) && (!blocks || S.key_nmeas == S.nmeas) &&
                            S.key_u.size() == nu && std::memcmp(S.key_u.data(), u_matrix, nu * sizeof(double)) == 0 &&
                            S.key_mu.size() == nmu && std::memcmp(S.key_mu.data(), muimp_full, nmu * sizeof(double)) == 0;
        if (revive) {
            S.eng = std::move(S.parked);
            Engine& E = *S.eng;
            const unsigned keep_stream = R.stream;
            R.naccpair = R.ntau; R.accpair_max = S.beta; R.freeze_window = 0;
            E.R = R; E.R.stream = keep_stream;
            h2d(const_cast<double*>(E.T.ftau), fd.data(), fd.size());
            E.set_taudiffmax(S.cfg_taudiffmax);
            E.rebuild(-1);
            E.zero_accum(); E.reset_counters();
        } else {
            S.parked.reset();
            AtomicTables at = build_atomic(nb, p.str("QuantumNumbers"), u_matrix, muimp_full, p.real("EPSBLOCK"), p.real("EPSEVEC"));
            S.eng.reset(new Engine());
            S.eng->symmoves = symmoves;
            S.eng->setup(at, fd, nftau, S.beta, R, opt);
            S.rho2_ready = false;
            S.key_u.assign(u_matrix, u_matrix + nu); S.key_mu.assign(muimp_full, muimp_full + nmu); S.key_par = key_of(p); S.key_opt = opt; S.key_nftau = nftau; S.key_nmeas = S.nmeas;
        }
        S.dens_off = S.eng->dens_off;                                        // block k: d x d column-major slab
        if (!revive && p.has("AccumGtau") && p.integer("AccumGtau") == 2) {  // autocorracc_create(ncomp, ilog2(NMeas))  CTQMC.F90:497-499
            long long v = S.nmeas, lv = -1; while (v > 0) { ++lv; v /= 2; }   // ilog2, Accumulators.F90:89-101
            S.eng->autocorr_setup((int)std::max<long long>(lv, 1));
        }
        S.have_res = false; S.time_warmup = S.time_sim = S.time_sim_steps = 0;
        S.aborted = 0; g_signal = 0;
        arm_traps();                                                         // CTQMC.F90:8076-8078
        S.eng->sync();
    });
}

int w2b_set_taudiffmax(double t) { return guarded([&] { require_engine(); S.eng->set_taudiffmax(t); }); }
int w2b_init_counters(void) { return guarded([&] { require_engine(); S.eng->reset_counters(); }); }

int w2b_ctqmc_warmup(long long nwarmups) {
    return guarded([&] {
        require_engine();
        double t0 = now();
        S.eng->set_accpair_binning(S.eng->R.ntau, S.beta);              // NAccPair = NGtau, AccPairMax = beta (CTQMC.F90:8203-8204)
        S.calib_accpair.clear();
        S.eng->reset_counters();
        const long long chunk = 1000;
        for (long long done = 0; done < nwarmups && !signal_fired(); done += chunk) S.eng->sweep(std::min(chunk, nwarmups - done));   // CTQMC.F90:8214
        S.eng->sync();
        if (nwarmups > 0) check_flags_early("during warm-up");
        S.time_warmup = now() - t0;
    });
}

int w2b_sample(long long nmeas, int reset) {
    return guarded([&] {
        require_engine();
        double t0 = now();
        if (reset) S.eng->zero_accum();
        S.eng->set_accpair_binning(S.eng->R.ntau, S.beta);              // CTQMC.F90:8391-8392
        S.calib_accpair.clear();
        S.eng->reset_counters();
        run_blocks(nmeas, true);
        S.time_sim += now() - t0;
    });
}
int w2b_finalize(void) { return guarded([&] { finalize_results(); }); }

int w2b_ctqmc_measure(int isector, int icomponent) {
    if (isector != 1) return fail("only the partition-function sector (isector = 1) is implemented");
    (void)icomponent;
    int rc = w2b_sample(S.nmeas, 1);
    if (rc) return rc;
    if (G.nranks > 1) { rc = w2b_allreduce_accum(); if (rc) return rc; }
    if (S.aborted == 2 && G.nranks == 1) return 0;               // no measurement was taken: nothing to normalise; the caller reads `aborted` (cthyb_solver.py:347)
    return w2b_finalize();
}

// ---- rank combination: one process per GPU, chains sharded by global chain id, ONE all-reduce of the raw
// accumulators per ctqmc_measure.  A host in any language drives it: rank 0 obtains the NCCL unique id, the host broadcasts
// those bytes by its own means (MPI_Bcast in the Fortran / mpi4py hosts, torch.distributed in bench.py), every rank calls
// w2b_comm_init BEFORE init_solver.  Replaces the per-quantity Allreduce of DistributedSample (statistics.py:95-132) for the sums;
// the caller's mean over ranks of the (now identical) results is then a no-op.
int w2b_comm_unique_id(void* out, size_t cap, size_t* n) {
    return guarded([&] {
#if W2B_DEVICE
        G.load();
        ncclUniqueId id; G.ck(G.GetUniqueId(&id), "ncclGetUniqueId");
        *n = sizeof id;
        if (out) { if (cap < sizeof id) throw std::runtime_error("w2b_comm_unique_id: buffer too small"); std::memcpy(out, &id, sizeof id); }
#else
        *n = 128; if (out) std::memset(out, 0, std::min<size_t>(cap, 128));
#endif
    });
}
int w2b_comm_init(const void* unique_id, size_t nbytes, int rank, int nranks) {
    return guarded([&] {
        if (nranks < 1 || rank < 0 || rank >= nranks) throw std::runtime_error("w2b_comm_init: bad rank / size");
#if W2B_DEVICE
        if (G.comm) { G.CommDestroy(G.comm); G.comm = nullptr; }
        if (nranks > 1) {
            G.load();
            ncclUniqueId id; if (nbytes != sizeof id) throw std::runtime_error("w2b_comm_init: unique id has the wrong size");
            std::memcpy(&id, unique_id, sizeof id);
            if (S.device >= 0) W2B_CUDA_CHECK(cudaSetDevice(S.device));
            G.ck(G.CommInitRank(&G.comm, nranks, id, rank), "ncclCommInitRank");
        }
#else
        (void)unique_id; (void)nbytes;
#endif
        G.rank = rank; G.nranks = nranks;
    });
}
int w2b_comm_destroy(void) {
    return guarded([&] {
#if W2B_DEVICE
        if (G.comm) { G.CommDestroy(G.comm); G.comm = nullptr; }
#endif
        G.rank = 0; G.nranks = 1;
    });
}
// sum of the raw accumulator slabs over all ranks, in place on the device (NCCL over NVLink / NVSwitch), and the maximum of the
// chains' error flags, so that every rank raises together.  No-op without a communicator.
int w2b_allreduce_accum(void) {
    return guarded([&] {
        require_engine();
        Engine& E = *S.eng;
        if (G.nranks <= 1) return;
        int flags[2] = {E.error_flags(), S.aborted};
        const size_t n = (size_t)E.R.nbins * E.accum_doubles;
#if W2B_DEVICE
        if (!G.comm) throw std::runtime_error("w2b_allreduce_accum: no communicator (call w2b_comm_init)");
        int* dfl = E.flag_scratch();
        h2d(dfl, flags, 2);
        G.ck(G.AllReduce(E.A.base, E.A.base, n, ncclDouble, ncclSum, G.comm, 0), "ncclAllReduce(accumulators)");
        G.ck(G.AllReduce(dfl, dfl, 2, ncclInt, ncclMax, G.comm, 0), "ncclAllReduce(flags)");
        W2B_CUDA_CHECK(cudaStreamSynchronize(0));
        d2h(flags, dfl, 2);
#else
        if (!G.hook) throw std::runtime_error("w2b_allreduce_accum: this build has no collective (tests install a hook)");
        G.hook(E.A.base, n, flags);
#endif
        if (flags[0]) E.force_error(flags[0]);                   // a flagged chain anywhere fails every rank's finalize
        if (flags[1] > S.aborted) S.aborted = flags[1] == 2 && S.nmeas_done > 0 ? 1 : flags[1];
    });
}
#if !W2B_DEVICE
void w2b_test_set_allreduce_hook(void (*fn)(double*, size_t, int*)) { G.hook = fn; }
#endif

// ctqmc_calibrate (CTQMC.F90:8088-8178), driven by CtHybSolver.calibrate_taudiff_max (cthyb_solver.py:899-921).
// Phase 1 (list): run until phase1pairnum pair moves were accepted; `accpairtau` = their tau distances, listed on the device
//   as they are accepted (the whole chain population contributes; the caller uses the 99.5 % quantile).
// Phase 2: histogram of accepted pair distances on [0, taudiff_max] with phase2taugrid bins, until every flavour holds
//   phase2pairnum entries; the caller fits an exponential.  As in the reference the window is never shifted here.
int w2b_ctqmc_calibrate(int list, int phase1pairnum, int phase2taugrid, int phase2pairnum, int nprogress) {
    (void)nprogress;
    return guarded([&] {
        require_engine();
        Engine& E = *S.eng;
        const int nf = E.T.nf;
        E.R.freeze_window = 1;
        std::vector<unsigned long long> ap;
        auto flavour_sums = [&](std::vector<unsigned long long>& sums) {
            E.get_accpair(ap); sums.assign(nf, 0ull);
            for (int f = 0; f < nf; ++f) for (int b = 0; b < E.R.naccpair; ++b) sums[f] += ap[(size_t)f * E.R.naccpair + b];
        };
        std::vector<unsigned long long> sums;
        try {
            if (list) {
                // AccPairTau(apt_index) = tau distance of every accepted pair move until phase1pairnum are listed (CTQMC.F90:3583-3586,
                // 4123-4126, 8159-8163); here the first phase1pairnum accepted by the whole population
                E.set_accpair_binning(E.R.ntau, S.beta);
                E.pairlist_begin(phase1pairnum);
                for (int it = 0; E.pairlist_count() < phase1pairnum && !signal_fired(); ++it) {
                    if (it == 100000) throw std::runtime_error("ctqmc_calibrate phase 1: no pair moves are being accepted");
                    E.sweep(50);
                }
                E.pairlist_get(S.accpairtau, std::min(phase1pairnum, E.pairlist_count()));
                S.accpairtau.resize(phase1pairnum, 0.0);
                E.pairlist_begin(0);
            } else {
                E.set_accpair_binning(phase2taugrid, E.R.taudiff_max);
                for (int it = 0;; ++it) {
                    E.sweep(50); flavour_sums(sums);
                    bool done = true; for (auto v : sums) if (v < (unsigned long long)phase2pairnum) done = false;
                    if (done) break;
                    if (it == 99999) throw std::runtime_error("ctqmc_calibrate phase 2: a flavour accepts no pair moves");
                }
                std::vector<double> out((size_t)nf * phase2taugrid);
                const int nbd = E.T.nb;
                for (int bb = 0; bb < nbd; ++bb) for (int s = 0; s < 2; ++s) for (int t = 0; t < phase2taugrid; ++t)
                    out[bb + (size_t)nbd * (s + 2 * t)] = (double)ap[(size_t)(2 * bb + s) * phase2taugrid + t];
                S.calib_accpair = out;
            }
        } catch (...) { E.R.freeze_window = 0; E.A.pairlist_cap = 0; throw; }
        E.R.freeze_window = 0;
        E.sync();
    });
}

int w2b_block(long long nsweep, int do_measure) {
    return guarded([&] { require_engine(); S.eng->sweep(nsweep); if (do_measure) { S.eng->measure(); S.eng->autocorr_add(); } });
}
int w2b_sync(void) { return guarded([&] { require_engine(); S.eng->sync(); }); }

// bench hook: nblocks x (sweep of nsweep moves per chain [+ one measurement]) timed with CUDA events on the
// launching stream.  ms[0] = whole region, ms[1] = sum of the sweep kernels, ms[2] = sum of the measurement kernels.
int w2b_bench_blocks(long long nblocks, long long nsweep, int do_measure, double* ms) {
    return guarded([&] {
        require_engine();
        Engine& E = *S.eng;
        ms[0] = ms[1] = ms[2] = 0.0;
#if W2B_DEVICE
        std::vector<cudaEvent_t> ev((size_t)nblocks * 2 + 1);
        for (auto& e : ev) W2B_CUDA_CHECK(cudaEventCreate(&e));
        E.sync();
        W2B_CUDA_CHECK(cudaEventRecord(ev[0]));
        for (long long i = 0; i < nblocks; ++i) {
            E.sweep(nsweep);
            W2B_CUDA_CHECK(cudaEventRecord(ev[2 * i + 1]));
            if (do_measure) E.measure();
            W2B_CUDA_CHECK(cudaEventRecord(ev[2 * i + 2]));
        }
        W2B_CUDA_CHECK(cudaEventSynchronize(ev[2 * nblocks]));
        float t = 0;
        W2B_CUDA_CHECK(cudaEventElapsedTime(&t, ev[0], ev[2 * nblocks])); ms[0] = t;
        for (long long i = 0; i < nblocks; ++i) {
            W2B_CUDA_CHECK(cudaEventElapsedTime(&t, ev[2 * i], ev[2 * i + 1])); ms[1] += t;
            W2B_CUDA_CHECK(cudaEventElapsedTime(&t, ev[2 * i + 1], ev[2 * i + 2])); ms[2] += t;
        }
        for (auto& e : ev) cudaEventDestroy(e);
#else
        double t0 = now();
        for (long long i = 0; i < nblocks; ++i) { double a = now(); E.sweep(nsweep); double b = now(); if (do_measure) E.measure(); double c = now(); ms[1] += (b - a) * 1e3; ms[2] += (c - b) * 1e3; }
        ms[0] = (now() - t0) * 1e3;
#endif
    });
}

// end-to-end hooks: new hybridisation function from host memory (what a DMFT iteration hands to init_solver,
// CTQMC.F90:7950-7955) and the raw accumulator slab back to host memory.
int w2b_upload_ftau(const double* ftau_full, int nftau) {
    return guarded([&] {
        require_engine();
        Engine& E = *S.eng;
        if (nftau != E.T.nftau) throw std::runtime_error("w2b_upload_ftau: Nftau differs from init_solver");
        const int nb = E.T.nb, nf = E.T.nf;
        std::vector<double>& fd = E.ftau_stage; fd.resize((size_t)nf * nftau);
        for (int b = 0; b < nb; ++b) for (int s = 0; s < 2; ++s) for (int t = 0; t < nftau; ++t)
            fd[(size_t)(2 * b + s) * nftau + t] = ftau_full[b + (size_t)nb * (s + 2 * (b + (size_t)nb * (s + 2 * (size_t)t)))];
        h2d(const_cast<double*>(E.T.ftau), fd.data(), fd.size());
    });
}
int w2b_read_accum(double* out, size_t cap, size_t* n) {
    return guarded([&] {
        require_engine();
        Engine& E = *S.eng;
        *n = (size_t)E.R.nbins * E.accum_doubles;
        if (out) d2h(out, E.A.base, std::min(cap, *n));
    });
}
void* w2b_pinned_alloc(size_t bytes) {
#if W2B_DEVICE
    void* p = nullptr; if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr; return p;
#else
    return std::malloc(bytes);
#endif
}
void w2b_pinned_free(void* p) {
#if W2B_DEVICE
    cudaFreeHost(p);
#else
    std::free(p);
#endif
}

int w2b_accum_device(void** p, size_t* n) {
    return guarded([&] { require_engine(); *p = (void*)S.eng->A.base; *n = (size_t)S.eng->R.nbins * S.eng->accum_doubles; });
}

int w2b_dest_ctqmc(void) {
    disarm_traps();
    if (S.reuse_population && S.eng) S.parked = std::move(S.eng);            // kept for the next init_solver of the same local problem
    else { S.eng.reset(); S.parked.reset(); }
    S.eng.reset(); S.have_res = false; S.raw_acc.clear(); S.bins_ready = false;
    return 0;
}
int w2b_dest_paras(void) { S.par.clear(); S.have_par = false; return 0; }

// ---------------------------------------------------------------- configurations
int w2b_get_mc_config_scalars(int chain, int* noper, int* outer_sst, int* outer_state) {
    return guarded([&] {
        check_chain(chain);
        std::vector<int> ih; std::vector<double> dh; S.eng->get_header(chain, ih, dh);
        *noper = ih[IH_NOPS]; *outer_sst = ih[IH_OUTER]; *outer_state = ih[IH_OSTATE] + 1;
    });
}
int w2b_get_mc_config(int chain, int n, double* taus, int* orbs, int* spins, int* cas, int* hashybs, int* outer_sst, int* outer_state) {
    return guarded([&] {
        check_chain(chain);
        std::vector<double> tau; std::vector<unsigned char> code; std::vector<short> sst;
        S.eng->get_ops(chain, tau, code, sst);
        if ((int)tau.size() != n) throw std::runtime_error("get_mc_config: size mismatch");
        for (int i = 0; i < n; ++i) { int f = code[i] >> 1; taus[i] = tau[i]; orbs[i] = f / 2 + 1; spins[i] = f % 2 + 1; cas[i] = (code[i] & 1) + 1; hashybs[i] = -1; }
        std::vector<int> ih; std::vector<double> dh; S.eng->get_header(chain, ih, dh);
        *outer_sst = ih[IH_OUTER]; *outer_state = ih[IH_OSTATE] + 1;
    });
}
int w2b_set_mc_config(int chain, int n, const double* taus, const int* orbs, const int* spins, const int* cas, const int* hashybs, int outer_sst, int outer_state) {
    return guarded([&] {
        check_chain(chain);
        const Engine& E = *S.eng;
        if (n < 0 || n > E.C.ops_cap) throw std::runtime_error("set_mc_config: " + std::to_string(n) + " operators exceed the capacity " + std::to_string(E.C.ops_cap) + " (option ops_cap)");
        if (outer_sst < 0 || outer_sst >= E.at.nsst) throw std::runtime_error("set_mc_config: outer superstate " + std::to_string(outer_sst) + " outside [0, " + std::to_string(E.at.nsst) + ")");
        // outer eigenstate (1-based like the reference, Trace.F90:6918): meaningful with state sampling only, else always the whole superstate
        if (E.R.statesampling && (outer_state < 1 || outer_state > E.at.dim[outer_sst])) throw std::runtime_error("set_mc_config: outer state " + std::to_string(outer_state) + " outside 1.." + std::to_string(E.at.dim[outer_sst]));
        std::vector<int> kf(E.at.nf, 0);
        std::vector<unsigned char> code(n);
        for (int i = 0; i < n; ++i) {
            if (orbs[i] < 1 || orbs[i] > E.at.nb || spins[i] < 1 || spins[i] > 2 || cas[i] < 1 || cas[i] > 2) throw std::runtime_error("set_mc_config: operator " + std::to_string(i) + " has orbital / spin / ca outside 1.." + std::to_string(E.at.nb) + " / 1..2 / 1..2");
            if (!(taus[i] >= 0.0 && taus[i] < S.beta)) throw std::runtime_error("set_mc_config: tau outside [0, beta)");
            if (cas[i] == 1 && ++kf[(orbs[i] - 1) * 2 + (spins[i] - 1)] > E.C.kcap) throw std::runtime_error("set_mc_config: flavour order exceeds kcap " + std::to_string(E.C.kcap));
            if (hashybs && hashybs[i] == 0) throw std::runtime_error("worm operators are outside the implemented hot path");
            code[i] = (unsigned char)((((orbs[i] - 1) * 2 + (spins[i] - 1)) << 1) | (cas[i] - 1));
            if (i && !(taus[i] > taus[i - 1])) throw std::runtime_error("set_mc_config: operators must be in ascending tau");
        }
        S.eng->set_ops(chain, n, taus, code.data(), outer_sst, E.R.statesampling ? outer_state - 1 : 0);
    });
}
int w2b_set_empty_mc_config(int chain, int outer_sst, int outer_state) { return w2b_set_mc_config(chain, 0, nullptr, nullptr, nullptr, nullptr, nullptr, outer_sst, outer_state); }

// ---------------------------------------------------------------- results
int w2b_get_array(const char* name, double* out, size_t cap, size_t* n) {
    return guarded([&] {
        std::string k(name);
        const std::vector<double>* src = nullptr; std::vector<double> tmp;
        auto bins = [&](const std::vector<std::vector<double>>& b) { tmp.clear(); for (auto& v : b) tmp.insert(tmp.end(), v.begin(), v.end()); src = &tmp; };
        if (k == "accpairtau") { src = &S.accpairtau; *n = src->size(); if (out) std::memcpy(out, src->data(), std::min(cap, src->size()) * sizeof(double)); return; }
        if (k == "accpair" && !S.calib_accpair.empty()) { src = &S.calib_accpair; *n = src->size(); if (out) std::memcpy(out, src->data(), std::min(cap, src->size()) * sizeof(double)); return; }
        if (!S.have_res) throw std::runtime_error("no results: call ctqmc_measure first");
        if (k.size() > 5 && k.compare(k.size() - 5, 5, "_bins") == 0) bins_materialise();
        if (k == "gtau_bins") bins(S.bin_gtau);
        else if (k == "gleg_bins") bins(S.bin_gleg);
        else if (k == "giw_bins") bins(S.bin_giw);
        else if (k == "occ_bins") bins(S.bin_occ);
        else if (k == "sign_bins") { tmp = S.bin_sign; src = &tmp; }
        else { auto it = S.res.arr.find(k); if (it == S.res.arr.end()) throw std::runtime_error("unknown result array " + k); src = &it->second; }
        *n = src->size();
        if (out) std::memcpy(out, src->data(), std::min(cap, src->size()) * sizeof(double));
    });
}

int w2b_get_scalar(const char* name, double* out) {
    return guarded([&] {
        std::string k(name);
        if (k == "time_warmup") { *out = S.time_warmup; return; }
        if (k == "time_sim") { *out = S.time_sim; return; }
        if (k == "time_sim_steps") { *out = S.time_sim_steps; return; }
        if (k == "aborted") { *out = S.aborted; return; }
        if (k == "h2d_bytes") { *out = (double)h2d_bytes(); return; }
        if (k == "d2h_bytes") { *out = (double)d2h_bytes(); return; }
        if (k == "nmeas_done") { *out = (double)S.nmeas_done; return; }
        if (k == "comm_rank") { *out = G.rank; return; }
        if (k == "comm_size") { *out = G.nranks; return; }
        if (S.eng) {
            Engine& E = *S.eng;
            if (k == "nchains") { *out = E.C.nchains; return; }
            if (k == "nbins") { *out = E.R.nbins; return; }
            if (k == "nsst") { *out = E.at.nsst; return; }
            if (k == "nstates") { *out = E.at.nstates; return; }
            if (k == "nstatesmax") { *out = E.at.dmax; return; }
            if (k == "kcap") { *out = E.C.kcap; return; }
            if (k == "ops_cap") { *out = E.C.ops_cap; return; }
            if (k == "num_windows") { *out = E.R.num_windows; return; }
            if (k == "nslide") { *out = E.R.nslide; return; }
            if (k == "naccpair") { *out = E.R.naccpair; return; }
            if (k == "taudiff_max") { *out = E.R.taudiff_max; return; }
            if (k == "egs") { *out = E.at.egs; return; }
            if (k == "gpu_launches") { *out = (double)E.launches; return; }
            if (k == "cta_warps") { *out = E.C.dmax == 1 ? E.ct_warps : E.sw_warps; return; }
            if (k == "batch_chains") { *out = E.ct_ncb; return; }
            if (k == "sweep_grid") { *out = E.C.dmax == 1 ? E.ct_grid : E.sw_grid; return; }
            if (k == "sweep_smem") { *out = (double)(E.C.dmax == 1 ? E.ct_smem : E.sw_smem); return; }
            if (k == "rounds") { *out = (double)E.rounds; return; }
            if (k == "trace_flops") { *out = (double)E.trace_flops(); return; }
            if (k == "ms_propose") { *out = E.ms_propose; return; }
            if (k == "ms_trace") { *out = E.ms_trace; return; }
            if (k == "ms_decide") { *out = E.ms_decide; return; }
            if (k == "blocked") { *out = E.C.dmax > 1 ? 1.0 : 0.0; return; }
            if (k.rfind("cnt_", 0) == 0) {
                static const char* names[C_NCNT] = {"try_add", "try_rem", "try_glob", "try_shift", "acc_add", "acc_rem", "acc_glob", "acc_shift",
                                                    "qn_add", "qn_rem", "qn_glob", "steps", "alg_bytes", "trace_flops", "try_add4", "acc_add4", "try_rem4", "acc_rem4", "four_rank2"};
                std::vector<unsigned long long> cnt; E.get_counters(cnt);
                for (int i = 0; i < C_NCNT; ++i) if (k.substr(4) == names[i]) { *out = (double)cnt[i]; return; }
            }
            if (k == "mean_order") { *out = E.mean_order(); return; }
            if (k.rfind("tick", 0) == 0 && k.size() == 5) {             // W2B_TIMING builds: clock cycles per move category of the autonomous sweep (diagnostics)
                unsigned long long v[8] = {0}; E.sync(); d2h((int*)v, E.C.sched + SC_STRIDE, 16); *out = (double)v[k[4] - '0']; return;
            }
            if (k == "max_k" || k == "max_ops" || k == "error_flags") { int mk, mo; int e = E.error_flags(&mk, &mo); *out = k == "max_k" ? mk : k == "max_ops" ? mo : e; return; }
        }
        if (S.have_res) { auto it = S.res.scal.find(k); if (it != S.res.scal.end()) { *out = it->second; return; } }
        throw std::runtime_error("unknown scalar " + k);
    });
}

// ---------------------------------------------------------------- replay / parity hooks
int w2b_debug_steps(int chain, long long nsteps, double* log) {
    return guarded([&] { check_chain(chain); S.eng->sweep(nsteps, chain, log); S.eng->sync(); });
}
int w2b_debug_chain_scalars(int chain, double* out) {
    return guarded([&] {
        check_chain(chain);
        std::vector<int> ih; std::vector<double> dh; S.eng->get_header(chain, ih, dh);
        out[0] = dh[DH_TRLOG]; out[1] = dh[DH_TRSIGN]; out[2] = dh[DH_DETLOG]; out[3] = dh[DH_DETSIGN]; out[4] = dh[DH_CFGSIGN];
        out[5] = ih[IH_NOPS]; out[6] = ih[IH_OUTER]; out[7] = ih[IH_ERR];
    });
}
int w2b_debug_minv(int chain, int orbital, int spin, double* out, size_t cap, int* k) {
    return guarded([&] {
        check_chain(chain);
        if (orbital < 1 || orbital > S.eng->at.nb || spin < 1 || spin > 2) throw std::runtime_error("debug_minv: orbital / spin out of range");
        std::vector<double> m; S.eng->get_minv(chain, (orbital - 1) * 2 + (spin - 1), m, *k);
        if (out) std::memcpy(out, m.data(), std::min(cap, m.size()) * sizeof(double));
    });
}

}  // extern "C"
