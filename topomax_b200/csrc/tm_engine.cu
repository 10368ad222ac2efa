// libtopomax_b200: engine (workspace, solvers, row-strip sharding) and the C ABI of
// include/topomax_b200.h.
//
// Sharding model: the mesh is cut into strips of cell rows, one per rank
// (one process per GPU).  A rank stores its owned lattice rows plus halo rows (2 cell rows
// below, 1 above) as one local lattice; kernels treat it as a standalone mesh and write only
// owned rows.  Halo rows are contiguous in memory, so an exchange is plain ncclSend/ncclRecv on
// the vector itself (4 lattice rows up, 3 down).  Dot products are reduced on the device and
// summed over ranks in place with ncclAllReduce -- no host round trip.  The top `dist_levels`
// multigrid levels are sharded; from there down every rank holds the (small) level in full and
// runs it redundantly, the hand-over being a gather by broadcasts of the restricted residual.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "../../include/topomax_b200.h"
#include "tm_comm.h"
#include "tm_dem.cuh"
#include "tm_elast.cuh"
#include "tm_filter_pcg.cuh"
#include "tm_fluid_cuda.cuh"
#include "tm_mg.cuh"
#include "tm_p1.cuh"
#include "tm_p1mg.cuh"
#include "tm_p2p.cuh"
#include "tm_tail.cuh"
#include "tm_vec.cuh"

namespace tmx {

static thread_local std::string g_error;
std::atomic<long long> g_launches{0};
void set_error(const std::string& msg) { g_error = msg; }
const char* last_error() { return g_error.c_str(); }

// Peer-memory halo exchange / all-reduce (tm_p2p.cuh) is the DEFAULT transport of sharded engines since
// round 2 (green on 2 and 4 real peers; 139 ms against 203 ms per step under NCCL on the 4-GPU BASELINE
// config, profiles/r2e_*); TM_P2P=0 or TM_OPT_P2P = 0 selects NCCL, which is also the automatic
// fall-back when any rank cannot map a peer's window.
static bool p2p_default() {
    const char* e = std::getenv("TM_P2P");
    return !(e && e[0] == '0');
}

struct Unsupported {
    std::string what;
};
struct Invalid {
    std::string what;
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void ensure(size_t count) {
        if (n >= count && p) return;
        release();
        TM_CUDA(cudaMalloc(&p, count * sizeof(T)));
        TM_CUDA(cudaMemset(p, 0, count * sizeof(T)));
        n = count;
    }
};

// peer-mapped windows of all ranks of the node (tm_p2p.cuh); shared with the fp32 twin engine
struct P2PState {
    char* win[P2P_MAX_RANKS] = {};
    size_t slot_bytes = 0;
    int rank = 0, nranks = 1;
    void close_peers() {
        for (int q = 0; q < nranks; ++q)
            if (q != rank && win[q]) {
                cudaIpcCloseMemHandle(win[q]);
                win[q] = nullptr;
            }
    }
    ~P2PState() {
        close_peers();
        if (win[rank]) cudaFree(win[rank]);
    }
};

struct SolveStats {
    int iters = 0;
    double relres = 0.0;
    bool converged = false;
    double floor = 0.0;      // estimate of the relative residual fp arithmetic cannot resolve (0: not estimated)
    double rtol_used = 0.0;  // max(rtol, floor_factor * floor): what the iteration was stopped at
};

// Measurement ledger (tm_ledger_read): every launch / collective of the engine files its
// ALGORITHMIC bytes (unique operand bytes read + written) under one category.
// Launches replayed from a CUDA graph are filed with the totals recorded at capture.  With
// TM_OPT_PROFILE = 3 every launch is also followed by an event record, and the time between two
// consecutive records is attributed to the category of the later one (one stream, in order: the
// figure includes the launch gap in front of a kernel, i.e. what the step really spends there).
enum LedgerCat {
    LC_FINE_PLAIN = 0, LC_FINE_DOT, LC_FINE_RESID, LC_FINE_CHEB, LC_FINE_RESID0, LC_FINE_CHEBDOT,
    LC_COARSE1 = 6,  // + (level - 1), levels 1 .. 14
    LC_RESTRICT = 20, LC_PROLONG, LC_CHEB_FIRST, LC_PCG_UPDATE, LC_PCG_DIRECTION, LC_REDUCTIONS, LC_COPIES,
    LC_TAIL, LC_COARSE_SOLVE, LC_FILTER, LC_MD, LC_SENS, LC_SETUP_COARSEN, LC_SETUP_DIAG, LC_SETUP_EIG,
    LC_HALO, LC_ALLREDUCE, LC_GATHER, LC_CONVERT, LC_MISC, LC_COUNT = 40
};
struct Ledger {
    double bytes[LC_COUNT], launches[LC_COUNT];
    Ledger() { clear(); }
    void clear() {
        for (int i = 0; i < LC_COUNT; ++i) bytes[i] = launches[i] = 0.0;
    }
    void add(const Ledger& o) {
        for (int i = 0; i < LC_COUNT; ++i) {
            bytes[i] += o.bytes[i];
            launches[i] += o.launches[i];
        }
    }
};

class EngineBase {
   public:
    virtual ~EngineBase() {}
    virtual void set_stream(cudaStream_t s) = 0;
    virtual void set_option(int opt, double value) = 0;
    virtual void comm_init(const char* id128) = 0;
    virtual void layout(int* out, int n) = 0;
    virtual void load_vector(const tm_loads& loads, void* b) = 0;
    virtual SolveStats filter_apply(int kind, void* in, void* out, double rtol, int maxit) = 0;
    virtual void elast_matvec(void* xi, double p, void* x, void* y) = 0;
    virtual void elast_diag(void* xi, double p, void* dinv) = 0;
    virtual SolveStats state_solve(void* xi, double p, const void* b, void* u, double rtol, int maxit,
                                   int flags) = 0;
    virtual double dot_p2(const void* u, const void* b) = 0;
    virtual void sens_rhs(const void* xi, double p, const void* u, void* out) = 0;
    virtual void md_halfstep(const void* psi, const void* g, double alpha, void* half) = 0;
    virtual void md_volume(const void* half, double c, double* vol, double* dvol) = 0;
    virtual void md_apply(const void* half, double c, const void* psi_prev, void* psi, void* rho,
                          double* delta_sq, double* vol) = 0;
    virtual int md_project(const void* half, double volume, double tol, int maxit, double* c, int* iters) = 0;
    virtual double integrate(const void* values) = 0;
    virtual void sample_field(int degree, const void* field, int nsx, int nsy, double x0, double dx, double y0,
                              double dy, void* out) = 0;
    virtual void last_stats(double* out, int n) = 0;
    virtual void mg_debug(void* xi, int op, int level, const void* in, void* out) = 0;
    virtual int mg_level_info(int level, int* info) = 0;
    virtual void profile_read(double* out, int n) = 0;
    virtual void ledger_read(double* out, int n, int reset) = 0;
    int device = 0;
};

// owned / stored cell-row ranges of one rank on one multigrid level
struct RowRange {
    int nyg;       // global cell rows of the level
    int c0, c1;    // owned cell rows [c0, c1)
    int cl0, cl1;  // stored cell rows [cl0, cl1)  (owned + halo)
    bool last;     // owns the top lattice row
    int ny() const { return cl1 - cl0; }
    int own_j0() const { return 2 * (c0 - cl0); }
    int own_j1() const { return 2 * (c1 - cl0) + (last ? 1 : 0); }
};

template <typename T>
class Engine : public EngineBase {
    template <typename U>
    friend class Engine;

   public:
    explicit Engine(const tm_config& cfg) : cfg_(cfg) {
        device = cfg.device;
        nx_ = cfg.nx;
        nyg_ = cfg.ny;
        rank_ = cfg.rank;
        nranks_ = std::max(1, cfg.nranks);
        if (nx_ < 1 || nyg_ < 1) throw Invalid{"nx, ny must be >= 1"};
        if (!(cfg.width > 0) || !(cfg.height > 0)) throw Invalid{"width/height must be > 0"};
        if (rank_ < 0 || rank_ >= nranks_) throw Invalid{"rank out of range"};
        hx_ = cfg.width / nx_;
        hy_ = cfg.height / nyg_;
        coarse_cells_ = 4;
        plan_levels(cfg.mg_dist_levels);
        derive_local_sizes();

        TM_CUDA(cudaSetDevice(device));
        TM_CUDA(cudaDeviceGetAttribute(&num_sms_, cudaDevAttrMultiProcessorCount, device));

        const Material<double> matd = make_material<double>(cfg.lame_lambda, cfg.lame_mu, hx_, hy_);
        diag_tab_ = make_diag_table(matd);
        tr_tab_ = make_transfer_table();
        co_tab_ = make_coarsen_table();

        rs_.capacity = 1 << 20;
        TM_CUDA(cudaMalloc(&rs_.partials, sizeof(double) * 2 * rs_.capacity));
        TM_CUDA(cudaMalloc(&rs_.counter, sizeof(unsigned int)));
        TM_CUDA(cudaMemset(rs_.counter, 0, sizeof(unsigned int)));
        TM_CUDA(cudaMalloc(&sc_, sizeof(double) * SC_COUNT));
        TM_CUDA(cudaMemset(sc_, 0, sizeof(double) * SC_COUNT));
        TM_CUDA(cudaMalloc(&eig_sc_, sizeof(double) * 64));
        TM_CUDA(cudaMallocHost(&h_sc_, sizeof(double) * 128));
        TM_CUDA(cudaStreamCreate(&own_stream_));
        stream_ = own_stream_;
    }

    ~Engine() override {
        cudaSetDevice(device);
        cudaStreamSynchronize(stream_);
        // graphs that captured NCCL operations must go before the communicator does
        if (graph_exec_) cudaGraphExecDestroy(graph_exec_);
        if (fgraph_exec_) cudaGraphExecDestroy(fgraph_exec_);
        if (setup_graph_exec_) cudaGraphExecDestroy(setup_graph_exec_);
        setup_graph_exec_ = nullptr;
        graph_exec_ = nullptr;
        fgraph_exec_ = nullptr;
        inner_.reset();
        p2p_release();
        if (comm_ && owns_comm_) nccl().CommDestroy(comm_);
        cudaFree(rs_.partials);
        cudaFree(rs_.counter);
        cudaFree(sc_);
        cudaFree(eig_sc_);
        cudaFree(filter_part_);
        cudaFreeHost(h_sc_);
        for (auto& p : prof_pending_) prof_free_.push_back(p.second);
        for (auto& e : prof_free_) {
            cudaEventDestroy(e.first);
            cudaEventDestroy(e.second);
        }
        for (auto& m : ev_marks_) cudaEventDestroy(m.second);
        for (auto& e : ev_free_) cudaEventDestroy(e);
        if (own_stream_) cudaStreamDestroy(own_stream_);
    }

    // The legacy default stream cannot be captured into a CUDA graph, so work addressed to it is
    // issued on an engine-owned BLOCKING stream instead: blocking streams order implicitly with
    // the legacy stream, so torch work on stream 0 before/after a call is still sequenced.
    void set_stream(cudaStream_t s) override { stream_ = s ? s : own_stream_; }

    void set_option(int opt, double value) override {
        switch (opt) {
            case TM_OPT_PRECOND: precond_ = (int)value; break;
            case TM_OPT_CHEB_DEGREE: cheb_degree_ = std::max(1, (int)value); break;
            case TM_OPT_CHECK_EVERY: check_every_ = std::max(0, (int)value); break;
            case TM_OPT_MG_COARSE_CELLS:
                if (nranks_ > 1) throw Invalid{"TM_OPT_MG_COARSE_CELLS cannot change on a sharded engine"};
                coarse_cells_ = std::min(4, std::max(1, (int)value));
                levels_.clear();
                plan_levels(0);
                derive_local_sizes();
                break;
            case 100: cheb_ratio_ = value; break;
            case 101: eig_safety_ = value; break;
            case TM_OPT_PROFILE:  // 1: level-0 kernel, 2: every level, 3: every launch by ledger category
                profile_ = (int)value;
                if (profile_ == 3 && ev_marks_.empty()) mark(-1);
                break;
            case 102: blocks_per_sm_target_ = std::max(1, (int)value); break;
            case 103: min_rows_per_strip_ = std::max(1, (int)value); break;
            case 104: filter_persistent_ = value != 0.0; break;
            case 109: coarse_degree_ = std::max(0, (int)value); break;
            case 110: use_graph_ = value != 0.0; graph_dirty_ = true; break;
            case 111: filter_cheb_ = value != 0.0; break;
            case 112: filter_mg_mode_ = (int)value; break;
            case 115: eig_first_its_ = std::max(1, (int)value); break;
            case 116: graph_sharded_ = value != 0.0; graph_dirty_ = true; break;
            case 117: mixed_ = value != 0.0; break;
            case 118: fuse_first_ = value != 0.0; graph_dirty_ = true; break;
            case 119: tail_max_nodes_ = (int)value; levels_.clear(); graph_dirty_ = true; break;
            case 122: tail_dry_ = (int)value; break;
            case 123: filter_tb_ = value != 0.0; break;
            case 125: warm_guard_ = value != 0.0; break;
            case 126: pdl_ = value != 0.0; graph_dirty_ = true; break;
            case 127: fuse_rz_ = (int)value; graph_dirty_ = true; break;
            case 128: restrict_tiled_ = value != 0.0; graph_dirty_ = true; break;
            case 129: wave_aware_ = value != 0.0; graph_dirty_ = true; break;
            case 131: fuse_cheb0_ = value != 0.0; graph_dirty_ = true; break;
            case 132: prolong_tiled_ = value != 0.0; graph_dirty_ = true; break;
            case 133:  // cycle window: first level whose visits are repeated (with 134, 135)
                cycle_first_ = std::max(1, (int)value); graph_dirty_ = true; break;
            case 134: cycle_last_ = (int)value; graph_dirty_ = true; break;
            case 136: light_levels_ = (int)value; graph_dirty_ = true; break;
            case 137: floor_factor_ = std::max(0.0, value); break;
            case 135:  // 0 = automatic window (default), 1 = V-cycle, 2.. = cycles per visit inside [133, 134]
                cycle_gamma_ = std::min(4, std::max(0, (int)value)); graph_dirty_ = true; break;
            case TM_OPT_P2P:  // collective: every rank must set it alike
                p2p_want_ = value != 0.0;
                graph_dirty_ = true;
                if (!p2p_want_) p2p_release();
                else if (comm_ && !p2p_) p2p_setup();
                break;
            case 124: filter_tb_steps_ = std::max(1, (int)value); filter_tb_state_ = 0; break;
            case 121: depth_limit_ = (int)value; graph_dirty_ = true; break;
            case 120: tail_cluster_ = std::max(1, (int)value); levels_.clear(); graph_dirty_ = true; break;
            case 113: filter_mg_degree_ = std::max(1, (int)value); break;
            case 114: filter_mg_ratio_ = value; break;
            case 105: apply_minb_ = std::min(5, std::max(2, (int)value)); break;
            case 106: apply_prefetch_ = value != 0.0; break;
            case 107:
                filter_blocks_per_sm_ = std::max(1, (int)value);
                if (filter_part_) cudaFree(filter_part_);
                filter_part_ = nullptr;
                filter_blocks_ = 0;
                break;
            default: throw Invalid{"unknown option " + std::to_string(opt)};
        }
    }

    // ------------------------------------------------------------------ sharding plumbing
    void comm_init(const char* id128) override {
        if (nranks_ == 1) return;
        if (!nccl().load()) throw Invalid{nccl().error};
        NcclUniqueId id;
        std::memcpy(id.internal, id128, 128);
        const int rc = nccl().CommInitRank(&comm_, nranks_, id, rank_);
        if (rc != 0) throw Invalid{std::string("ncclCommInitRank: ") + nccl().GetErrorString(rc)};
        if (p2p_want_) p2p_setup();
    }

    // Peer windows (tm_p2p.cuh): allocate mine, all-gather the IPC handles through NCCL, map the
    // others.  Collective; if ANY rank fails to map a peer, every rank stays on the NCCL path.
    void p2p_setup() {
        need_comm();
        if (nranks_ > P2P_MAX_RANKS) return;
        auto st = std::make_shared<P2PState>();
        st->rank = rank_;
        st->nranks = nranks_;
        // largest halo message: 4 lattice rows of a level-0 displacement vector
        st->slot_bytes = (((size_t)4 * (2 * (size_t)nx_ + 1) * 2 * sizeof(double)) + 255) & ~(size_t)255;
        // whole 2 MiB pages: an IPC handle exports the allocation's pages, and two windows of one
        // process (two engines alive at once) must never share a page
        const size_t page = (size_t)2 << 20;
        const size_t bytes = std::max(2 * page, (P2P_HEADER_BYTES + 4 * st->slot_bytes + page - 1) / page * page);
        int ok = 1;
        char* self = nullptr;
        cudaIpcMemHandle_t mine;
        std::memset(&mine, 0, sizeof(mine));
        if (cudaMalloc(&self, bytes) != cudaSuccess) {
            ok = 0;
        } else {
            st->win[rank_] = self;
            TM_CUDA(cudaMemsetAsync(self, 0, bytes, stream_));
            if (cudaIpcGetMemHandle(&mine, self) != cudaSuccess) ok = 0;
        }
        cudaGetLastError();
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        DevBuf<char> hb;
        hb.ensure((size_t)64 * nranks_);
        TM_CUDA(cudaMemcpyAsync(hb.p + 64 * rank_, &mine, 64, cudaMemcpyHostToDevice, stream_));
        nccl_check(nccl().GroupStart(), "ncclGroupStart");
        for (int q = 0; q < nranks_; ++q)
            nccl().Broadcast(hb.p + 64 * q, hb.p + 64 * q, 64, kNcclInt8, q, comm_, stream_);
        nccl_check(nccl().GroupEnd(), "ncclGroupEnd");
        std::vector<cudaIpcMemHandle_t> all(nranks_);
        TM_CUDA(cudaMemcpyAsync(all.data(), hb.p, (size_t)64 * nranks_, cudaMemcpyDeviceToHost, stream_));
        TM_CUDA(cudaStreamSynchronize(stream_));
        for (int q = 0; q < nranks_ && ok; ++q) {
            if (q == rank_) continue;
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                ok = 0;
                cudaGetLastError();
            } else {
                st->win[q] = static_cast<char*>(ptr);
            }
        }
        // agree: sum of the failure counts over ranks must be zero
        double* flag = sc_ + SC_TMP;
        const double bad = ok ? 0.0 : 1.0;
        TM_CUDA(cudaMemcpyAsync(flag, &bad, sizeof(double), cudaMemcpyHostToDevice, stream_));
        nccl_check(nccl().AllReduce(flag, flag, 1, kNcclFloat64, kNcclSum, comm_, stream_), "ncclAllReduce");
        double total = 1.0;
        TM_CUDA(cudaMemcpyAsync(&total, flag, sizeof(double), cudaMemcpyDeviceToHost, stream_));
        TM_CUDA(cudaStreamSynchronize(stream_));
        if (total == 0.0) {
            p2p_ = st;
            if (inner_) inner_->p2p_ = p2p_;
        }
        graph_dirty_ = true;
    }
    bool p2p_active() const { return static_cast<bool>(p2p_); }
    // Collective (the owner of the communicator only): every rank unmaps its peers' windows, all
    // ranks meet, and only then is the own window freed -- an exported allocation must outlive its
    // importers' mappings.
    void p2p_release() {
        if (!p2p_) return;
        if (owns_comm_ && comm_) {
            if (inner_) inner_->p2p_.reset();
            cudaStreamSynchronize(stream_);
            p2p_->close_peers();
            double* flag = sc_ + SC_TMP;
            if (nccl().AllReduce(flag, flag, 1, kNcclFloat64, kNcclSum, comm_, stream_) == 0)
                cudaStreamSynchronize(stream_);
        }
        p2p_.reset();
        graph_dirty_ = true;
    }

    void layout(int* out, int n) override {
        const RowRange& r = ranges_[0];
        const int v[11] = {rank_, nranks_, nx_, nyg_, r.cl0, r.cl1, r.c0, r.c1, r.last ? 1 : 0, dist_levels_,
                           p2p_active() ? 1 : 0};
        for (int i = 0; i < n && i < 11; ++i) out[i] = v[i];
    }

    // ------------------------------------------------------------------ loads
    void load_vector(const tm_loads& in, void* b) override {
        LoadSpec s;
        std::memset(&s, 0, sizeof(s));
        s.nx = nx_; s.ny = nyg_; s.W = cfg_.width; s.H = cfg_.height;
        s.j_off = g0_.j_off; s.Ly_loc = g0_.Ly;
        s.has_force = in.has_force;
        s.fcx = in.force_center[0]; s.fcy = in.force_center[1]; s.frad = in.force_radius;
        s.fx = in.force_value[0]; s.fy = in.force_value[1];
        if (in.ntractions < 0 || in.ntractions > 8) throw Invalid{"ntractions must be in 0..8"};
        // the reference compares the node coordinate with the side coordinate exactly
        // (FEM_src/elasisity_problem.py:54-66); ((n*W)/n == W) can fail for odd W
        const bool right_ok = ((double)nx_ * cfg_.width) / (double)nx_ == cfg_.width;
        const bool top_ok = ((double)nyg_ * cfg_.height) / (double)nyg_ == cfg_.height;
        for (int t = 0; t < in.ntractions; ++t) {
            int side;
            switch (in.traction_side[t]) {
                case TM_SIDE_LEFT: side = 0; break;
                case TM_SIDE_RIGHT: side = 1; break;
                case TM_SIDE_TOP: side = 2; break;
                case TM_SIDE_BOTTOM: side = 3; break;
                default: throw Invalid{"Malformed side: " + std::to_string(in.traction_side[t])};
            }
            if ((side == 1 && !right_ok) || (side == 2 && !top_ok)) continue;
            const int k = s.ntractions++;
            s.tside[k] = side;
            s.tlo[k] = in.traction_center[t] - in.traction_length[t] / 2;
            s.thi[k] = in.traction_center[t] + in.traction_length[t] / 2;
            s.tx[k] = in.traction_value[t][0];
            s.ty[k] = in.traction_value[t][1];
        }
        // exact P2 mass matrix of a triangle of area |T| (vertices 0..2, mids 01,12,02)
        const double area = 0.5 * hx_ * hy_;
        const int opposite_mid[3] = {4, 5, 3};  // mid(12) opposite v0, mid(02) opp v1, mid(01) opp v2
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) {
                double v;
                if (i < 3 && j < 3) v = (i == j) ? 1.0 / 30 : -1.0 / 180;
                else if (i >= 3 && j >= 3) v = (i == j) ? 8.0 / 45 : 4.0 / 45;
                else {
                    const int vert = i < 3 ? i : j, mid = i < 3 ? j : i;
                    v = (opposite_mid[vert] == mid) ? -1.0 / 45 : 0.0;
                }
                s.M2[i][j] = area * v;
            }
        dim3 blk(32, 8), grd(ceil_div(g0_.Lx, 32), ceil_div(g0_.Ly, 8));
        load_vector_kernel<T><<<grd, blk, 0, stream_>>>(s, (T*)b);
        TM_CHECK_LAUNCH();
        acct(LC_MISC, sz(nu_));
    }

    // ------------------------------------------------------------------ filter
    SolveStats filter_apply(int kind, void* in_, void* out_, double rtol, int maxit) override {
        T* in = (T*)in_;
        T* out = (T*)out_;
        const double alpha = cfg_.filter_radius * cfg_.filter_radius, beta = 1.0;
        if (kind != 0 && kind != 1) throw Invalid{"rhs_kind must be 0 or 1"};
        f_r_.ensure(n1_); f_p_.ensure(n1_); f_Ap_.ensure(n1_); f_rhs_.ensure(n1_);
        if (!f_dinv_ready_) {
            f_dinv_.ensure(n1_);
            p1_diag_kernel<T><<<grid2d_p1(), dim3(32, 8), 0, stream_>>>(p1_, alpha, beta, f_dinv_.p);
            TM_CHECK_LAUNCH();
            acct(LC_FILTER, sz(n1_));
            f_dinv_ready_ = true;
        }
        const bool use_fmg = filter_mg_mode_ == 1 ||
                             (filter_mg_mode_ == 0 && nlevels_ >= 2 && (nranks_ > 1 || n1_ > (size_t)2000000));
        if (use_fmg && nlevels_ >= 2) return filter_apply_mg(kind, in, out, alpha, beta, rtol, maxit);
        if (filter_persistent_ && nranks_ == 1) {
            // whole solve in one cooperative launch (tm_filter_pcg.cuh)
            const T* rhs_p = in;
            if (kind == 0) {
                p1_apply(0.0, 1.0, in, f_rhs_.p, nullptr);  // rhs = M1 in
                rhs_p = f_rhs_.p;
                if (out != in) {
                    TM_CUDA(cudaMemcpyAsync(out, in, n1_ * sizeof(T), cudaMemcpyDeviceToDevice, stream_));
                    acct(LC_FILTER, 2 * sz(n1_));
                }
            } else {
                TM_CUDA(cudaMemsetAsync(out, 0, n1_ * sizeof(T), stream_));
                acct(LC_FILTER, sz(n1_));
            }
            f_p2_.ensure(n1_);
            if (filter_blocks_ == 0) {
                int per_sm = 0;
                TM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, filter_pcg_kernel<T>,
                                                                      kFilterThreads, 0));
                if (per_sm < 1) throw Invalid{"filter_pcg_kernel cannot be made resident"};
                filter_blocks_ = std::min(per_sm, filter_blocks_per_sm_) * num_sms_;
                TM_CUDA(cudaMalloc(&filter_part_, sizeof(double) * (4 * (size_t)filter_blocks_ + 4)));
            }
            const int nb = (int)std::max<size_t>(
                1, std::min<size_t>(filter_blocks_, (n1_ + kFilterThreads - 1) / kFilterThreads));
            double* result = filter_part_ + 4 * (size_t)filter_blocks_;
            const T* dinv_p = f_dinv_.p;
            T *x_p = out, *r_p = f_r_.p, *Ap_p = f_Ap_.p, *p0_p = f_p_.p, *p1_p = f_p2_.p;
            SolveStats st;
            bool done = false;
            if (filter_cheb_ && filter_bounds_ready_) {
                // Chebyshev semi-iteration: one grid barrier per iteration, no dot products
                FilterChebArgs ca;
                ca.g = p1_; ca.alpha = alpha; ca.beta = beta;
                ca.lmin = filter_lmin_; ca.lmax = filter_lmax_; ca.rtol = rtol;
                ca.maxit = std::min(maxit, 4 * filter_cg_iters_ + 50); ca.check = 8;
                ca.part = filter_part_; ca.result = result;
                T* xalt_p = f_Ap_.p;
                T* d_p = f_p_.p;
                FilterTbArgs ta;
                if (filter_tb_ && plan_filter_tb(ta)) {
                    // temporally blocked: filter_tb_steps_ iterations per grid barrier
                    ta.lmin = filter_lmin_; ta.lmax = filter_lmax_; ta.rtol = rtol;
                    ta.maxit = ca.maxit; ta.part = filter_part_; ta.result = result;
                    for (int cy = 0; cy < 3; ++cy)
                        for (int cx = 0; cx < 3; ++cx) p1_class_stencil(p1_, alpha, beta, cx, cy, ta.coef[3 * cy + cx]);
                    T* dalt_p = f_r_.p;
                    const double* c12_p = filter_c12(ta.maxit + ta.S + 1);
                    void* kargs[] = {&ta, &rhs_p, &c12_p, &x_p, &xalt_p, &d_p, &dalt_p};
                    TM_CUDA(cudaLaunchCooperativeKernel(filter_tb_entry(ta.rows_per_thread),
                                                        dim3(ta.tiles_x * ta.tiles_y), dim3(kFilterTbThreads), kargs,
                                                        filter_tb_smem_, stream_));
                } else {
                    void* kargs[] = {&ca, &rhs_p, &dinv_p, &x_p, &xalt_p, &d_p};
                    TM_CUDA(cudaLaunchCooperativeKernel((void*)filter_cheb_kernel<T>, dim3(nb),
                                                        dim3(kFilterThreads), kargs, 0, stream_));
                }
                ++g_launches;
                TM_CUDA(cudaMemcpyAsync(h_sc_ + 64, result, sizeof(double) * 3, cudaMemcpyDeviceToHost, stream_));
                TM_CUDA(cudaStreamSynchronize(stream_));
                st.iters = (int)h_sc_[64];
                st.relres = h_sc_[65];
                st.converged = h_sc_[66] != 0.0;
                // one Chebyshev-Jacobi iteration: read x, b, D^-1, d; write x, d
                acct(LC_FILTER, 6.0 * st.iters * sz(n1_));
                done = st.converged;  // otherwise continue with CG from the current iterate,
                if (!done) filter_bounds_ready_ = false;  // ... and re-estimate the spectral bounds
            }
            if (!done) {
                FilterPcgArgs fa;
                fa.g = p1_; fa.alpha = alpha; fa.beta = beta; fa.rtol = rtol; fa.maxit = maxit;
                fa.partA = filter_part_;
                fa.partB = filter_part_ + filter_blocks_;
                fa.result = result;
                const bool record = filter_cheb_ && !filter_bounds_ready_;
                if (record) filter_coef_.ensure(2 * 4096);
                fa.coef = record ? filter_coef_.p : nullptr;
                fa.coef_cap = 4096;
                void* kargs[] = {&fa, &rhs_p, &dinv_p, &x_p, &r_p, &Ap_p, &p0_p, &p1_p};
                TM_CUDA(cudaLaunchCooperativeKernel((void*)filter_pcg_kernel<T>, dim3(nb), dim3(kFilterThreads),
                                                    kargs, 0, stream_));
                ++g_launches;
                TM_CUDA(cudaMemcpyAsync(h_sc_ + 64, result, sizeof(double) * 3, cudaMemcpyDeviceToHost, stream_));
                TM_CUDA(cudaStreamSynchronize(stream_));
                const int cg_its = (int)h_sc_[64];
                acct(LC_FILTER, 13.0 * cg_its * sz(n1_));  // one Jacobi-PCG iteration = 13 n1
                st.iters += cg_its;
                st.relres = h_sc_[65];
                st.converged = h_sc_[66] != 0.0;
                if (record && st.converged && cg_its >= 8) lanczos_bounds(std::min(cg_its, 4096));
            }
            return st;
        }
        // launch-per-operation PCG (sharded runs: the search direction needs a halo exchange)
        const T* rhs;
        if (kind == 0) {
            exchange_p1(in);
            p1_apply(0.0, 1.0, in, f_rhs_.p, nullptr);  // rhs = M1 in
            rhs = f_rhs_.p;
            if (out != in) {
                TM_CUDA(cudaMemcpyAsync(out, in, n1_ * sizeof(T), cudaMemcpyDeviceToDevice, stream_));
                acct(LC_FILTER, 2 * sz(n1_));
            }
            p1_apply(alpha, beta, out, f_Ap_.p, nullptr);
            waxpby_kernel<T><<<grid1d(p1_cnt_), kVecThreads, 0, stream_>>>(
                p1_cnt_, 1.0, rhs + p1_off_, -1.0, f_Ap_.p + p1_off_, f_r_.p + p1_off_);
            TM_CHECK_LAUNCH();
            acct(LC_FILTER, 3 * sz(p1_cnt_));
        } else {
            rhs = in;
            TM_CUDA(cudaMemsetAsync(out, 0, n1_ * sizeof(T), stream_));
            TM_CUDA(cudaMemcpyAsync(f_r_.p, rhs, n1_ * sizeof(T), cudaMemcpyDeviceToDevice, stream_));
            acct(LC_FILTER, 3 * sz(n1_));
        }
        auto apply_dot = [&](T* p, T* Ap) {
            exchange_p1(p);
            p1_apply(alpha, beta, p, Ap, sc_ + SC_PAP);
            sum_ranks(sc_ + SC_PAP, 1);
        };
        auto precond = [&](T*) -> T* { return nullptr; };
        const int check = check_every_ > 0 ? check_every_ : 10;
        SolveStats st = pcg(p1_off_, p1_cnt_, rhs, out, f_r_.p, f_p_.p, f_Ap_.p, f_dinv_.p, apply_dot,
                            precond, true, rtol, maxit, check, true);
        exchange_p1(out);
        return st;
    }

    // ------------------------------------------------------------------ filter multigrid
    struct FLevel {
        P1Level<T> L;      // local geometry + stencils
        P1Geom gpiece;     // first replicated level: full geometry, own rows = my piece
        size_t n = 0, off = 0, cnt = 0;
        bool sharded = false;
        DevBuf<T> S, dinv, x, xalt, b, d, tmp, eig;
        double lmax = 0.0;
    };

    P1Geom p1_level_geom(int l, bool replicated) const {
        const RowRange r = row_range(l, rank_, replicated);
        P1Geom g = p1_;
        g.nx = lv_nx_[l];
        g.ny = r.ny();
        g.hx = hx_ * (1 << l);
        g.hy = hy_ * (1 << l);
        g.iy_off = r.cl0;
        g.ny_global = r.nyg;
        g.own_iy0 = r.c0 - r.cl0;
        g.own_iy1 = r.c1 - r.cl0 + (r.last ? 1 : 0);
        return g;
    }
    // halo exchange of a P1 array of sharded level l (2 vertex rows each way)
    void exchange_p1_level(int l, T* v) {
        if (nranks_ == 1 || l >= dist_levels_) return;
        const RowRange& r = ranges_[l];
        exchange_rows(v, (size_t)lv_nx_[l] + 1, r.c0 - r.cl0, r.c1 - r.cl0, 2, 2, r.last);
    }
    // gather of a FULL P1 array of level l from the owned vertex rows of every rank
    void gather_p1_rows(int l, T* full) {
        if (nranks_ == 1) return;
        need_comm();
        const int dtype = sizeof(T) == 8 ? kNcclFloat64 : kNcclFloat32;
        const size_t row_elems = (size_t)lv_nx_[l] + 1;
        nccl_check(nccl().GroupStart(), "ncclGroupStart");
        for (int q = 0; q < nranks_; ++q) {
            const int j0 = starts_[q] >> l;
            const int j1 = q == nranks_ - 1 ? lv_ny_[l] + 1 : (starts_[q + 1] >> l);
            if (j1 <= j0) continue;
            T* ptr = full + (size_t)j0 * row_elems;
            nccl().Broadcast(ptr, ptr, (size_t)(j1 - j0) * row_elems, dtype, q, comm_, stream_);
        }
        nccl_check(nccl().GroupEnd(), "ncclGroupEnd");
        acct(LC_GATHER, sz(((size_t)lv_ny_[l] + 1) * row_elems));
    }

    void fmg_apply(int l, int ep, const T* x, T* y, const T* b, double c1, double c2, double* dot_out) {
        FLevel& F = flevels_[l];
        dim3 blk(32, 8), grd(ceil_div(F.L.g.nx + 1, 32), ceil_div(F.L.g.ny + 1, 8));
        if ((long)grd.x * grd.y > rs_.capacity) throw Invalid{"reduction scratch too small"};
        switch (ep) {
            case P1EP_PLAIN:
                p1mg_apply_kernel<T, P1EP_PLAIN><<<grd, blk, 0, stream_>>>(F.L, x, y, b, F.dinv.p, F.d.p, c1, c2, rs_, dot_out);
                break;
            case P1EP_DOT:
                p1mg_apply_kernel<T, P1EP_DOT><<<grd, blk, 0, stream_>>>(F.L, x, y, b, F.dinv.p, F.d.p, c1, c2, rs_, dot_out);
                break;
            case P1EP_RESID:
                p1mg_apply_kernel<T, P1EP_RESID><<<grd, blk, 0, stream_>>>(F.L, x, y, b, F.dinv.p, F.d.p, c1, c2, rs_, dot_out);
                break;
            default:
                p1mg_apply_kernel<T, P1EP_CHEB><<<grd, blk, 0, stream_>>>(F.L, x, y, b, F.dinv.p, F.d.p, c1, c2, rs_, dot_out);
                break;
        }
        TM_CHECK_LAUNCH();
        {   // operands: x, y (+ b; + D^-1, d read and written) + the 7 stencil planes of a coarse level
            const double vecs = ep == P1EP_PLAIN || ep == P1EP_DOT ? 2.0 : (ep == P1EP_RESID ? 3.0 : 6.0);
            acct(LC_FILTER, (vecs + (l > 0 ? 7.0 : 0.0)) * sz(F.n));
        }
    }

    // one-off: Galerkin stencils, diagonals, smoother bounds, coarse factor (operator is constant)
    void build_filter_mg(double alpha, double beta) {
        flevels_.clear();
        flevels_.resize(nlevels_);
        for (int l = 0; l < nlevels_; ++l) {
            FLevel& F = flevels_[l];
            F.sharded = l < dist_levels_;
            F.L.g = p1_level_geom(l, !F.sharded);
            F.L.alpha = alpha;
            F.L.beta = beta;
            F.L.S = nullptr;
            F.gpiece = F.L.g;
            const size_t W1 = (size_t)F.L.g.nx + 1;
            F.n = W1 * (F.L.g.ny + 1);
            F.off = (size_t)F.L.g.own_iy0 * W1;
            F.cnt = (size_t)(F.L.g.own_iy1 - F.L.g.own_iy0) * W1;
            if (nranks_ > 1 && l == dist_levels_) {
                F.gpiece.own_iy0 = starts_[rank_] >> l;
                F.gpiece.own_iy1 = rank_ == nranks_ - 1 ? lv_ny_[l] + 1 : (starts_[rank_ + 1] >> l);
            }
            const bool coarsest = l + 1 == nlevels_;
            F.x.ensure(F.n);
            if (l > 0) {
                F.S.ensure(7 * F.n);
                F.b.ensure(F.n);
            }
            if (!coarsest) {
                F.dinv.ensure(F.n); F.xalt.ensure(F.n); F.d.ensure(F.n); F.tmp.ensure(F.n); F.eig.ensure(F.n);
            }
            if (F.n > (size_t)kP1CoarseMax && coarsest) throw Invalid{"coarsest filter level too large"};
        }
        dim3 blk(32, 8);
        for (int l = 1; l < nlevels_; ++l) {
            FLevel& Ff = flevels_[l - 1];
            FLevel& Fc = flevels_[l];
            const bool gather = nranks_ > 1 && l == dist_levels_;
            const int c_off = Fc.sharded ? ranges_[l].cl0 : 0;
            int own0 = 0, own1 = Fc.L.g.ny + 1;
            if (Fc.sharded || gather) {
                const RowRange rc = row_range(l, rank_, false);
                own0 = rc.c0 - c_off;
                own1 = rc.c1 - c_off + (rc.last ? 1 : 0);
            }
            dim3 grd(ceil_div(Fc.L.g.nx + 1, 32), ceil_div(Fc.L.g.ny + 1, 8));
            p1mg_coarsen_kernel<T><<<grd, blk, 0, stream_>>>(Ff.L, lv_ny_[l - 1], Ff.L.g.iy_off, Fc.L.g.nx, Fc.L.g.ny,
                                                            c_off, own0, own1, Fc.S.p);
            TM_CHECK_LAUNCH();
            for (int k = 0; k < 7; ++k) {
                if (Fc.sharded) exchange_p1_level(l, Fc.S.p + k * Fc.n);
                if (gather) gather_p1_rows(l, Fc.S.p + k * Fc.n);
            }
            Fc.L.S = Fc.S.p;
        }
        for (int l = 0; l + 1 < nlevels_; ++l) {
            FLevel& F = flevels_[l];
            dim3 grd(ceil_div(F.L.g.nx + 1, 32), ceil_div(F.L.g.ny + 1, 8));
            p1mg_diag_kernel<T><<<grd, blk, 0, stream_>>>(F.L, F.dinv.p);
            TM_CHECK_LAUNCH();
            // lambda_max(D^-1 A) by power iteration from a constant+ramp start
            const int g1 = grid1d(F.cnt);
            waxpby_kernel<T><<<grid1d(F.n), kVecThreads, 0, stream_>>>(F.n, 0.0, F.dinv.p, 0.0, F.dinv.p, F.eig.p);
            TM_CHECK_LAUNCH();
            fill_alternating_kernel<T><<<grid1d(F.n), kVecThreads, 0, stream_>>>(F.n, F.eig.p);
            TM_CHECK_LAUNCH();
            double* slot = eig_sc_ + 32 + l;
            dot_kernel<T><<<g1, kVecThreads, 0, stream_>>>(F.cnt, F.eig.p + F.off, F.eig.p + F.off, rs_, slot);
            TM_CHECK_LAUNCH();
            if (F.sharded) sum_ranks(slot, 1);
            for (int it = 0; it < 15; ++it) {
                exchange_p1_level(l, F.eig.p);
                fmg_apply(l, P1EP_PLAIN, F.eig.p, F.tmp.p, nullptr, 0, 0, nullptr);
                normalize_scale_kernel<T><<<g1, kVecThreads, 0, stream_>>>(F.cnt, F.dinv.p + F.off, F.tmp.p + F.off,
                                                                          F.eig.p + F.off, slot);
                TM_CHECK_LAUNCH();
                dot_kernel<T><<<g1, kVecThreads, 0, stream_>>>(F.cnt, F.eig.p + F.off, F.eig.p + F.off, rs_, slot);
                TM_CHECK_LAUNCH();
                if (F.sharded) sum_ranks(slot, 1);
            }
        }
        {
            FLevel& C = flevels_[nlevels_ - 1];
            fcoarse_A_.ensure(C.n * C.n);
            p1mg_coarse_factor_kernel<T><<<1, 64, 0, stream_>>>(C.L, fcoarse_A_.p);
            TM_CHECK_LAUNCH();
        }
        TM_CUDA(cudaMemcpyAsync(h_sc_ + 80, eig_sc_ + 32, sizeof(double) * 32, cudaMemcpyDeviceToHost, stream_));
        TM_CUDA(cudaStreamSynchronize(stream_));
        for (int l = 0; l + 1 < nlevels_; ++l) {
            const double lam = std::sqrt(h_sc_[80 + l]);
            if (!(lam > 0.0) || !(lam == lam)) throw Invalid{"filter multigrid: eigenvalue estimate failed"};
            flevels_[l].lmax = lam;
        }
        fmg_ready_ = true;
    }

    T* fsmooth(int l, const T* b, T* xin) {
        FLevel& F = flevels_[l];
        const double hi = 1.1 * F.lmax, lo = hi / filter_mg_ratio_;
        const double theta = 0.5 * (hi + lo), delta = 0.5 * (hi - lo), sigma = theta / delta;
        double rho = 1.0 / sigma;
        T* cur;
        int k0 = 0;
        if (!xin) {
            cheb_first_kernel<T><<<grid1d(F.cnt), kVecThreads, 0, stream_>>>(
                F.cnt, 1.0 / theta, F.dinv.p + F.off, b + F.off, F.d.p + F.off, F.x.p + F.off);
            TM_CHECK_LAUNCH();
            acct(LC_FILTER, 4 * sz(F.cnt));
            cur = F.x.p;
            k0 = 1;
        } else {
            cur = xin;
        }
        for (int k = k0; k < filter_mg_degree_; ++k) {
            double c1, c2;
            if (k == 0) {
                c1 = 0.0;
                c2 = 1.0 / theta;
            } else {
                const double rho_new = 1.0 / (2.0 * sigma - rho);
                c1 = rho_new * rho;
                c2 = 2.0 * rho_new / delta;
                rho = rho_new;
            }
            T* other = (cur == F.x.p) ? F.xalt.p : F.x.p;
            exchange_p1_level(l, cur);
            fmg_apply(l, P1EP_CHEB, cur, other, b, c1, c2, nullptr);
            cur = other;
        }
        return cur;
    }

    T* fvcycle_body(T* r) {
        const int nl = nlevels_;
        vcycle_rz_ = false;
        std::vector<T*> xs(nl, nullptr);
        std::vector<const T*> bs(nl, nullptr);
        bs[0] = r;
        dim3 blk(32, 8);
        for (int l = 0; l + 1 < nl; ++l) {
            FLevel& F = flevels_[l];
            FLevel& C = flevels_[l + 1];
            xs[l] = fsmooth(l, bs[l], nullptr);
            exchange_p1_level(l, xs[l]);
            fmg_apply(l, P1EP_RESID, xs[l], F.tmp.p, bs[l], 0, 0, nullptr);
            exchange_p1_level(l, F.tmp.p);
            const bool gather = nranks_ > 1 && (l + 1) == dist_levels_;
            dim3 grd(ceil_div(C.L.g.nx + 1, 32), ceil_div(C.L.g.ny + 1, 8));
            p1mg_restrict_kernel<T><<<grd, blk, 0, stream_>>>(F.L.g, lv_ny_[l], gather ? C.gpiece : C.L.g, F.tmp.p, C.b.p);
            TM_CHECK_LAUNCH();
            acct(LC_FILTER, sz(F.n) + sz(C.n));
            if (gather) gather_p1_rows(l + 1, C.b.p);
            bs[l + 1] = C.b.p;
        }
        {
            FLevel& C = flevels_[nl - 1];
            p1mg_coarse_solve_kernel<T><<<1, 32, 0, stream_>>>((int)C.n, fcoarse_A_.p, bs[nl - 1], C.x.p);
            TM_CHECK_LAUNCH();
            acct(LC_FILTER, 2 * sz(C.n));
            xs[nl - 1] = C.x.p;
        }
        for (int l = nl - 1; l-- > 0;) {
            FLevel& F = flevels_[l];
            FLevel& C = flevels_[l + 1];
            exchange_p1_level(l + 1, xs[l + 1]);
            dim3 grd(ceil_div(F.L.g.nx + 1, 32), ceil_div(F.L.g.ny + 1, 8));
            p1mg_prolong_add_kernel<T><<<grd, blk, 0, stream_>>>(F.L.g, C.L.g, xs[l + 1], xs[l]);
            TM_CHECK_LAUNCH();
            acct(LC_FILTER, 2 * sz(F.n) + sz(C.n));
            xs[l] = fsmooth(l, bs[l], xs[l]);
        }
        return xs[0];
    }

    // graph replay of the filter V-cycle (the operator never changes: captured once per engine)
    T* fvcycle(T* r) {
        if (!use_graph_ || (nranks_ > 1 && !graph_sharded_) || profile_ >= 2) return fvcycle_body(r);
        if (fgraph_exec_ && fgraph_r_ == r) {
            TM_CUDA(cudaGraphLaunch(fgraph_exec_, stream_));
            ++g_launches;
            led_.add(fgraph_led_);
            return fgraph_z_;
        }
        if (fgraph_exec_) {
            cudaGraphExecDestroy(fgraph_exec_);
            fgraph_exec_ = nullptr;
        }
        const long long l0 = g_launches.load();
        cudaGraph_t graph = nullptr;
        TM_CUDA(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal));
        CaptureLedger cap(led_, capturing_);
        T* z = nullptr;
        try {
            z = fvcycle_body(r);
        } catch (...) {
            cap.finish();
            cudaStreamEndCapture(stream_, &graph);
            if (graph) cudaGraphDestroy(graph);
            throw;
        }
        fgraph_led_ = cap.finish();
        TM_CUDA(cudaStreamEndCapture(stream_, &graph));
        g_launches.store(l0);
        TM_CUDA(cudaGraphInstantiate(&fgraph_exec_, graph, 0));
        cudaGraphDestroy(graph);
        fgraph_r_ = r;
        fgraph_z_ = z;
        return fvcycle(r);
    }

    SolveStats filter_apply_mg(int kind, T* in, T* out, double alpha, double beta, double rtol, int maxit) {
        if (!fmg_ready_) build_filter_mg(alpha, beta);
        const T* rhs;
        if (kind == 0) {
            exchange_p1(in);
            p1_apply(0.0, 1.0, in, f_rhs_.p, nullptr);  // rhs = M1 in
            rhs = f_rhs_.p;
            if (out != in) {
                TM_CUDA(cudaMemcpyAsync(out, in, n1_ * sizeof(T), cudaMemcpyDeviceToDevice, stream_));
                acct(LC_FILTER, 2 * sz(n1_));
            }
            p1_apply(alpha, beta, out, f_Ap_.p, nullptr);
            waxpby_kernel<T><<<grid1d(p1_cnt_), kVecThreads, 0, stream_>>>(
                p1_cnt_, 1.0, rhs + p1_off_, -1.0, f_Ap_.p + p1_off_, f_r_.p + p1_off_);
            TM_CHECK_LAUNCH();
            acct(LC_FILTER, 3 * sz(p1_cnt_));
        } else {
            rhs = in;
            TM_CUDA(cudaMemsetAsync(out, 0, n1_ * sizeof(T), stream_));
            TM_CUDA(cudaMemcpyAsync(f_r_.p, rhs, n1_ * sizeof(T), cudaMemcpyDeviceToDevice, stream_));
            acct(LC_FILTER, 3 * sz(n1_));
        }
        auto apply_dot = [&](T* p, T* Ap) {
            exchange_p1(p);
            p1_apply(alpha, beta, p, Ap, sc_ + SC_PAP);
            sum_ranks(sc_ + SC_PAP, 1);
        };
        auto precond = [&](T* r) -> T* { return fvcycle(r); };
        SolveStats st = pcg(p1_off_, p1_cnt_, rhs, out, f_r_.p, f_p_.p, f_Ap_.p, nullptr, apply_dot, precond, false,
                            rtol, maxit, 1, true);
        exchange_p1(out);
        return st;
    }

    // Extreme eigenvalues of D^-1 A_f from the Lanczos tridiagonal of one converged CG solve
    // (diag_k = 1/a_k + b_{k-1}/a_{k-1}, off_k = sqrt(b_k)/a_k), by Sturm bisection.
    void lanczos_bounds(int m) {
        std::vector<double> c(2 * m);
        TM_CUDA(cudaMemcpyAsync(c.data(), filter_coef_.p, sizeof(double) * 2 * m, cudaMemcpyDeviceToHost, stream_));
        TM_CUDA(cudaStreamSynchronize(stream_));
        std::vector<double> d(m), e(m, 0.0);
        for (int k = 0; k < m; ++k) {
            const double ak = c[2 * k], bk = c[2 * k + 1];
            if (!(ak > 0.0)) return;
            d[k] = 1.0 / ak + (k > 0 ? c[2 * (k - 1) + 1] / c[2 * (k - 1)] : 0.0);
            e[k] = std::sqrt(std::max(bk, 0.0)) / ak;  // couples k and k+1
        }
        auto count_below = [&](double x) {
            int cnt = 0;
            double q = d[0] - x;
            if (q < 0) ++cnt;
            for (int k = 1; k < m; ++k) {
                if (q == 0.0) q = 1e-300;
                q = d[k] - x - e[k - 1] * e[k - 1] / q;
                if (q < 0) ++cnt;
            }
            return cnt;
        };
        double lo = 0.0, hi = 0.0;
        for (int k = 0; k < m; ++k) hi = std::max(hi, d[k] + (k > 0 ? e[k - 1] : 0.0) + e[k]);
        auto kth = [&](int k) {  // k-th smallest eigenvalue
            double a = lo, b = hi;
            for (int it = 0; it < 200; ++it) {
                const double mid = 0.5 * (a + b);
                if (count_below(mid) > k) b = mid; else a = mid;
            }
            return 0.5 * (a + b);
        };
        const double emin = kth(0), emax = kth(m - 1);
        if (!(emin > 0.0) || !(emax > emin)) return;
        // Ritz values converge from inside the spectrum; keep the widest interval seen so far
        filter_lmin_ = filter_lmin_ > 0.0 ? std::min(filter_lmin_, 0.9 * emin) : 0.9 * emin;
        filter_lmax_ = std::max(filter_lmax_, 1.05 * emax);
        filter_cg_iters_ = m;
        filter_bounds_ready_ = true;
    }

    // ------------------------------------------------------------------ elasticity operator
    // SIMP exponent of the current call.  p = 3 (every elasticity design of the reference) keeps
    // the closed-form moments inside the fine-level operator; any other p > 0 evaluates level-0
    // moments once per call into w0_ (mg_fine_moments_kernel) and level 0 runs the stored-moment
    // kernels, like the coarse levels (reference: src/penalizers.py:36-46, the penalties loop of
    // src/solver.py:230-231).
    void set_penalty(double p) {
        if (!(p > 0.0) || !(p < 1e6)) throw Invalid{"penalty p=" + std::to_string(p) + ": p must be positive"};
        if (p == spec_.p) return;
        spec_.p = p;
        const int ip = (int)p;
        spec_.ip = ((double)ip == p && ip >= 1 && ip <= kMaxIntPenalty) ? ip : 0;
        general_ = p != 3.0;
        // captured kernels carry the coefficient source by value; smoother bounds restart
        if (setup_graph_exec_) {
            cudaGraphExecDestroy(setup_graph_exec_);
            setup_graph_exec_ = nullptr;
        }
        graph_dirty_ = true;
        for (auto& L : levels_) L.eig_ready = false;
    }
    void compute_w0(const T* xi) {
        if (!general_) return;
        w0_.ensure((size_t)12 * g0_.nx * g0_.ny);
        LevelGeom<T> g = g0_;
        g.xi = xi;
        dim3 blk(32, 8), grd(ceil_div(g.nx, 32), ceil_div(g.ny, 8));
        mg_fine_moments_kernel<T><<<grd, blk, 0, stream_>>>(g, spec_, w0_.p);
        TM_CHECK_LAUNCH();
        acct(LC_SETUP_COARSEN, sz(n1_) + 12 * sz((size_t)g.nx * g.ny));
    }
    // level-0 geometry bound to a density (its halo rows already exchanged)
    LevelGeom<T> fine_geom(const T* xi, bool compute) {
        LevelGeom<T> g = g0_;
        g.xi = xi;
        if (general_) {
            w0_.ensure((size_t)12 * g0_.nx * g0_.ny);
            if (compute) compute_w0(xi);
            g.W = w0_.p;
        }
        return g;
    }

    void elast_matvec(void* xi, double p, void* x, void* y) override {
        set_penalty(p);
        if (x == y) throw Invalid{"tm_elast_matvec: x and y must not alias"};
        exchange_p1((T*)xi);
        exchange_p2(0, (T*)x);
        const LevelGeom<T> g = fine_geom((const T*)xi, true);
        ApplyArgs<T> a = apply_args();
        a.x = (const T*)x;
        a.y = (T*)y;
        launch_apply(g, false, EP_PLAIN, a);
    }

    void elast_diag(void* xi, double p, void* dinv) override {
        set_penalty(p);
        exchange_p1((T*)xi);
        const LevelGeom<T> g = fine_geom((const T*)xi, true);
        launch_diag(g, false, (T*)dinv);
    }

    SolveStats state_solve(void* xi_, double p, const void* b_, void* u_, double rtol, int maxit,
                           int flags) override {
        set_penalty(p);
        T* xi = (T*)xi_;
        const T* b = (const T*)b_;
        T* u = (T*)u_;
        s_r_.ensure(nu_); s_p_.ensure(nu_); s_Ap_.ensure(nu_); s_b_.ensure(nu_);
        stats_fine_applies_ = 0;
        stats_vcycles_ = 0;

        exchange_p1(xi);
        // engine-owned copy: the captured set-up / V-cycle graphs stay valid whatever buffer the
        // caller's allocator hands over from one solve to the next
        s_xi_.ensure(n1_);
        TM_CUDA(cudaMemcpyAsync(s_xi_.p, xi, n1_ * sizeof(T), cudaMemcpyDeviceToDevice, stream_));
        acct(LC_COPIES, 2 * sz(n1_));
        xi = s_xi_.p;
        LevelGeom<T> g = g0_;
        g.xi = xi;
        mask_fixed_kernel<T><<<grid1d(n2_), kVecThreads, 0, stream_>>>(g, b, s_b_.p);
        TM_CHECK_LAUNCH();
        acct(LC_COPIES, 2 * sz(nu_));

        bool use_mg = precond_ == TM_PRECOND_MULTIGRID && nlevels_ >= 2;
        // mixed precision: the multigrid preconditioner runs in fp32 inside the fp64 PCG (the
        // converged solution and its residual test stay fp64)
        const bool mixed = use_mg && mixed_ && sizeof(T) == 8;
        const T* dinv = nullptr;
        // general exponent: the level-0 moments come from setup_device_part when this engine owns
        // the hierarchy, else from here
        g = fine_geom(xi, !(use_mg && !mixed));
        if (use_mg && mixed) {
            prepare_inner(xi);
        } else if (use_mg) {
            if (levels_.empty()) build_levels();
            setup_hierarchy(xi);
        } else {
            s_dinv_.ensure(nu_);
            launch_diag(g, false, s_dinv_.p);
            dinv = s_dinv_.p;
        }

        bool warm = (flags & 1) != 0;
        stats_warm_used_ = 0;
        if (warm) {
            mask_fixed_kernel<T><<<grid1d(n2_), kVecThreads, 0, stream_>>>(g, u, u);
            TM_CHECK_LAUNCH();
            acct(LC_COPIES, 2 * sz(nu_));
            exchange_p2(0, u);
            ApplyArgs<T> a = apply_args();
            a.x = u; a.y = s_r_.p; a.b = s_b_.p;
            launch_apply(g, false, EP_RESID, a);
            if (warm_guard_) {
                // The previous displacement is only a good guess once the design settles: while the
                // SIMP stiffness of whole regions still moves by orders of magnitude its residual can
                // exceed ||b|| (the residual of the zero guess) by far.  Keep whichever is smaller.
                const int g1 = grid1d(p2_cnt_);
                dot_kernel<T><<<g1, kVecThreads, 0, stream_>>>(p2_cnt_, s_r_.p + p2_off_, s_r_.p + p2_off_, rs_,
                                                              sc_ + SC_TMP);
                dot_kernel<T><<<g1, kVecThreads, 0, stream_>>>(p2_cnt_, s_b_.p + p2_off_, s_b_.p + p2_off_, rs_,
                                                              sc_ + SC_TMP + 1);
                TM_CHECK_LAUNCH();
                ++g_launches;
                acct(LC_REDUCTIONS, sz(p2_cnt_));
                acct(LC_REDUCTIONS, sz(p2_cnt_));
                sum_ranks(sc_ + SC_TMP, 2);
                read_scalars();
                if (!(h_sc_[SC_TMP] < h_sc_[SC_TMP + 1])) warm = false;
            }
            stats_warm_used_ = warm ? 1 : 0;
        }
        if (!warm) {
            TM_CUDA(cudaMemsetAsync(u, 0, nu_ * sizeof(T), stream_));
            TM_CUDA(cudaMemcpyAsync(s_r_.p, s_b_.p, nu_ * sizeof(T), cudaMemcpyDeviceToDevice, stream_));
            acct(LC_COPIES, 3 * sz(nu_));
        }
        // Option 137 (floor_factor_, default 0.5; 0 = off): with an initial guess at hand -- the previous
        // design's displacement, within a few per cent of the new one -- estimate the relative residual that
        // fp arithmetic cannot resolve, ||K (u o delta)|| / ||b|| with delta_i = +-eps/2, by ONE operator pass,
        // and stop the PCG at max(rtol, floor_factor x that).  On 2e8 dofs the level is 2.6e-9, on 6.4e9 dofs
        // 2.9e-6 (bench.py measures the same quantity with the independent CPU operator): a tolerance of 1e-10
        // is then met only by the recursive residual, the true one stays where it was several iterations earlier.
        double floor_factor = 0.0;
        if (warm && floor_factor_ > 0.0 && sizeof(T) == 8) {  // (the fp32 engine is a separately stated mode)
            half_ulp_kernel<T><<<grid1d(nu_), kVecThreads, 0, stream_>>>(nu_, (size_t)0, (const T*)u, s_p_.p);
            TM_CHECK_LAUNCH();
            acct(LC_MISC, 2 * sz(nu_));
            exchange_p2(0, s_p_.p);  // signs are drawn per local index: make the halo rows the owners' values
            ApplyArgs<T> a = apply_args();
            a.x = s_p_.p; a.y = s_Ap_.p;
            launch_apply(g, false, EP_PLAIN, a);
            dot_kernel<T><<<grid1d(p2_cnt_), kVecThreads, 0, stream_>>>(p2_cnt_, s_Ap_.p + p2_off_, s_Ap_.p + p2_off_, rs_,
                                                                        sc_ + SC_FLOOR);
            TM_CHECK_LAUNCH();
            acct(LC_REDUCTIONS, sz(p2_cnt_));
            sum_ranks(sc_ + SC_FLOOR, 1);
            floor_factor = floor_factor_;
        }

        auto apply_dot = [&](T* pp, T* Ap) {
            exchange_p2(0, pp);
            ApplyArgs<T> a = apply_args();
            a.x = pp; a.y = Ap; a.dot_out = sc_ + SC_PAP;
            launch_apply(g, false, EP_DOT, a);
            sum_ranks(sc_ + SC_PAP, 1);
        };
        auto precond = [&](T* r) -> T* { return !use_mg ? nullptr : (mixed ? mixed_vcycle(r) : vcycle(r)); };
        int check = check_every_;
        // automatic: with the multigrid preconditioner every iteration is checked, but only from where the
        // previous solve of this engine stopped (consecutive designs of a run need about the same count; the
        // host then enqueues the earlier iterations without waiting for the device)
        int check_from = 0;
        if (check <= 0) {
            check = use_mg ? 1 : 25;
            if (use_mg && warm && expected_iters_ > 3) check_from = expected_iters_ - 2;
        }
        SolveStats st = pcg(p2_off_, p2_cnt_, s_b_.p, u, s_r_.p, s_p_.p, s_Ap_.p, dinv, apply_dot, precond,
                            !use_mg, rtol, maxit, check, false, check_from, floor_factor);
        expected_iters_ = st.converged ? st.iters : 0;
        stats_floor_ = st.floor;
        stats_rtol_used_ = st.rtol_used;
        exchange_p2(0, u);
        stats_iters_ = st.iters;
        if (mixed) {
            stats_fine_applies_ += inner_->stats_fine_applies_;
            stats_vcycles_ = inner_->stats_vcycles_;
        }
        return st;
    }

    double dot_p2(const void* u, const void* b) override {
        dot_kernel<T><<<grid1d(p2_cnt_), kVecThreads, 0, stream_>>>(
            p2_cnt_, (const T*)u + p2_off_, (const T*)b + p2_off_, rs_, sc_ + SC_TMP);
        TM_CHECK_LAUNCH();
        acct(LC_REDUCTIONS, 2 * sz(p2_cnt_));
        sum_ranks(sc_ + SC_TMP, 1);
        read_scalars();
        return h_sc_[SC_TMP];
    }

    void sens_rhs(const void* xi, double p, const void* u, void* out) override {
        set_penalty(p);
        LevelGeom<T> g = g0_;
        g.xi = (const T*)xi;
        if (general_)
            sens_rhs_kernel<T, true><<<grid2d_p1(), dim3(32, 8), 0, stream_>>>(g, spec_, p1_.own_iy0, p1_.own_iy1,
                                                                              (const T*)u, (T*)out);
        else
            sens_rhs_kernel<T, false><<<grid2d_p1(), dim3(32, 8), 0, stream_>>>(g, spec_, p1_.own_iy0,
                                                                               p1_.own_iy1, (const T*)u, (T*)out);
        TM_CHECK_LAUNCH();
        acct(LC_SENS, sz(nu_) + 2 * sz(n1_));
    }

    // ------------------------------------------------------------------ mirror descent
    void md_halfstep(const void* psi, const void* g, double alpha, void* half) override {
        waxpby_kernel<T><<<grid1d(n1_), kVecThreads, 0, stream_>>>(n1_, 1.0, (const T*)psi, -alpha,
                                                                  (const T*)g, (T*)half);
        TM_CHECK_LAUNCH();
        acct(LC_MD, 3 * sz(n1_));
    }

    void md_volume(const void* half, double c, double* vol, double* dvol) override {
        md_volume_kernel<T><<<grid1d(n1_), kVecThreads, 0, stream_>>>(p1_, (const T*)half, c, rs_,
                                                                     sc_ + SC_TMP);
        TM_CHECK_LAUNCH();
        acct(LC_MD, sz(n1_));
        sum_ranks(sc_ + SC_TMP, 2);
        read_scalars();
        *vol = h_sc_[SC_TMP];
        *dvol = h_sc_[SC_TMP + 1];
    }

    // Newton iteration of the volume projection with the iterate resident on the device: batches of
    // `kBatch` (reduction, [sum over ranks,] scalar update) triples are enqueued blind and the state is
    // read back once per batch -- 2-3 host synchronisations per projection instead of two per iterate.
    // Returns 1 converged, 2 zero derivative, 0 not converged within maxit (the caller falls back to
    // Brent, src/solver.py:175-186).  Same IEEE operations as the host-driven loop: same iterates.
    int md_project(const void* half, double volume, double tol, int maxit, double* c, int* iters) override {
        constexpr int kBatch = 3;
        double* st = sc_ + SC_NEWTON;
        TM_CUDA(cudaMemsetAsync(st, 0, 4 * sizeof(double), stream_));
        int done_its = 0;
        while (true) {
            const int n = std::min(kBatch, maxit - done_its);
            for (int k = 0; k < n; ++k) {
                md_volume_kernel<T><<<grid1d(n1_), kVecThreads, 0, stream_>>>(p1_, (const T*)half, 0.0, rs_,
                                                                             sc_ + SC_TMP, st);
                TM_CHECK_LAUNCH();
                acct(LC_MD, sz(n1_));
                sum_ranks(sc_ + SC_TMP, 2);
                md_newton_update_kernel<<<1, 32, 0, stream_>>>(sc_ + SC_TMP, volume, tol, st);
                TM_CHECK_LAUNCH();
                acct(LC_MD, 0.0);
            }
            done_its += n;
            read_scalars();
            const double* h = h_sc_ + SC_NEWTON;
            if (h[1] != 0.0 || done_its >= maxit) {
                *c = h[0];
                *iters = (int)h[2];
                return h[1] != 0.0 ? (int)h[3] : 0;
            }
        }
    }

    void md_apply(const void* half, double c, const void* psi_prev, void* psi, void* rho,
                  double* delta_sq, double* vol) override {
        md_apply_kernel<T><<<grid1d(n1_), kVecThreads, 0, stream_>>>(
            p1_, (const T*)half, c, (const T*)psi_prev, (T*)psi, (T*)rho, rs_, sc_ + SC_TMP);
        TM_CHECK_LAUNCH();
        acct(LC_MD, 4 * sz(n1_));
        sum_ranks(sc_ + SC_TMP, 2);
        read_scalars();
        *delta_sq = h_sc_[SC_TMP];
        *vol = h_sc_[SC_TMP + 1];
    }

    double integrate(const void* values) override {
        p1_integrate_kernel<T><<<grid1d(n1_), kVecThreads, 0, stream_>>>(p1_, (const T*)values, rs_,
                                                                        sc_ + SC_TMP);
        TM_CHECK_LAUNCH();
        acct(LC_MD, sz(n1_));
        sum_ranks(sc_ + SC_TMP, 1);
        read_scalars();
        return h_sc_[SC_TMP];
    }

    void sample_field(int degree, const void* field, int nsx, int nsy, double x0, double dx, double y0,
                      double dy, void* out) override {
        if (nranks_ > 1) throw Unsupported{"tm_sample_field works on an unsharded engine (gather the field first)"};
        if (degree != 1 && degree != 2) throw Invalid{"tm_sample_field: degree must be 1 (P1) or 2 (vector P2)"};
        if (nsx < 0 || nsy < 0) throw Invalid{"tm_sample_field: negative sample count"};
        if (nsx == 0 || nsy == 0) return;
        dim3 blk(32, 8), grd(ceil_div(nsx, 32), ceil_div(nsy, 8));
        if (degree == 1)
            sample_field_kernel<T, 1><<<grd, blk, 0, stream_>>>(nx_, nyg_, hx_, hy_, (const T*)field, nsx, nsy, x0,
                                                               dx, y0, dy, (T*)out);
        else
            sample_field_kernel<T, 2><<<grd, blk, 0, stream_>>>(nx_, nyg_, hx_, hy_, (const T*)field, nsx, nsy, x0,
                                                               dx, y0, dy, (T*)out);
        TM_CHECK_LAUNCH();
    }

    void last_stats(double* out, int n) override {
        // [5..8]: cumulative level-0 operator launches per epilogue (plain, dot, residual, Chebyshev)
        long epc[4];
        for (int e = 0; e < 4; ++e) epc[e] = fine_ep_count_[e] + (inner_ ? inner_->fine_ep_count_[e] : 0);
        const double lmax0 = inner_ && !inner_->levels_.empty() ? inner_->levels_[0].lmax
                                                                : (levels_.empty() ? 0.0 : levels_[0].lmax);
        // [9]: first level of the cluster tail (-1: none), [10]: its cluster size
        const int tf = inner_ ? inner_->tail_first_ : tail_first_;
        const int tc = inner_ ? inner_->tail_cluster_used_ : tail_cluster_used_;
        // [15]: relative fp floor estimated for the last state solve (0: none), [16]: the tolerance it stopped at
        // [12..14]: the levels cycled more than once per visit of their parent (first, last, cycles; -1 -1 1: V-cycle)
        int cf = -1, cl = -1, cg = 1;
        for (int l = 1; l + 1 < nlevels_; ++l)
            if (repeats(l) > 1) {
                if (cf < 0) cf = l;
                cl = l;
                cg = repeats(l);
            }
        const double v[17] = {(double)stats_iters_, (double)stats_vcycles_, (double)stats_fine_applies_,
                              (double)nlevels_, lmax0, (double)epc[0], (double)epc[1], (double)epc[2],
                              (double)epc[3], (double)tf, (double)(tf >= 0 ? tc : 0), (double)stats_warm_used_,
                              (double)cf, (double)cl, (double)cg, stats_floor_, stats_rtol_used_};
        for (int i = 0; i < n && i < 17; ++i) out[i] = v[i];
    }

    // CUDA-event timing of every fine-level operator launch (TM_OPT_PROFILE), per epilogue
    // out[0..3] ms / out[4..7] launches of the level-0 kernel per epilogue; then for every level
    // l < 16: out[8+8l .. +3] ms and out[8+8l+4 .. +7] launches of that level's operator kernel
    void profile_read(double* out, int n) override {
        std::vector<double> ms(64, 0.0), cnt(64, 0.0);
        TM_CUDA(cudaStreamSynchronize(stream_));
        for (auto& p : prof_pending_) {
            float t = 0.f;
            TM_CUDA(cudaEventElapsedTime(&t, p.second.first, p.second.second));
            if (p.first < 64) {
                ms[p.first] += t;
                cnt[p.first] += 1;
            }
            prof_free_.push_back(p.second);
        }
        prof_pending_.clear();
        for (int i = 0; i < n; ++i) out[i] = 0.0;
        for (int i = 0; i < 4; ++i) {
            if (i < n) out[i] = ms[i];
            if (4 + i < n) out[4 + i] = cnt[i];
        }
        for (int l = 0; l < 16; ++l)
            for (int e = 0; e < 4; ++e) {
                if (8 + 8 * l + e < n) out[8 + 8 * l + e] = ms[4 * l + e];
                if (8 + 8 * l + 4 + e < n) out[8 + 8 * l + 4 + e] = cnt[4 * l + e];
            }
        if (inner_) {  // the fp32 twin's launches (mixed precision) count as well
            std::vector<double> t(n, 0.0);
            inner_->profile_read(t.data(), n);
            for (int i = 0; i < n; ++i) out[i] += t[i];
        }
    }

    // out[0 .. LC_COUNT) algorithmic bytes, [LC_COUNT .. 2 LC_COUNT) launches, [2 LC_COUNT .. 3 LC_COUNT)
    // milliseconds (TM_OPT_PROFILE = 3 only) per LedgerCat since the last reset
    void ledger_read(double* out, int n, int reset) override {
        TM_CUDA(cudaStreamSynchronize(stream_));
        double ms[LC_COUNT];
        for (int i = 0; i < LC_COUNT; ++i) ms[i] = 0.0;
        for (size_t k = 1; k < ev_marks_.size(); ++k) {
            float t = 0.f;
            TM_CUDA(cudaEventElapsedTime(&t, ev_marks_[k - 1].second, ev_marks_[k].second));
            if (ev_marks_[k].first >= 0) ms[ev_marks_[k].first] += t;
        }
        for (int i = 0; i < n; ++i) out[i] = 0.0;
        for (int i = 0; i < LC_COUNT; ++i) {
            if (i < n) out[i] = led_.bytes[i];
            if (LC_COUNT + i < n) out[LC_COUNT + i] = led_.launches[i];
            if (2 * LC_COUNT + i < n) out[2 * LC_COUNT + i] = ms[i];
        }
        if (inner_) {  // the fp32 twin (mixed precision) files its own launches
            std::vector<double> t(n, 0.0);
            inner_->ledger_read(t.data(), n, reset);
            for (int i = 0; i < n; ++i) out[i] += t[i];
        }
        if (reset) {
            led_.clear();
            for (auto& m : ev_marks_) ev_free_.push_back(m.second);
            ev_marks_.clear();
            if (profile_ == 3) mark(-1);
        }
    }

    // diagnostics (tests): multigrid internals on the hierarchy built for xi (single rank)
    void mg_debug(void* xi, int op, int level, const void* in_, void* out_) override {
        if (nranks_ > 1) throw Invalid{"tm_mg_debug is a single-rank diagnostic"};
        if (nlevels_ < 2) throw Invalid{"no multigrid hierarchy for this mesh"};
        if (levels_.empty()) build_levels();
        const int nl = nlevels_;
        if (level < 0 || level >= nl) throw Invalid{"bad level"};
        setup_hierarchy((T*)xi);
        const T* in = (const T*)in_;
        T* out = (T*)out_;
        Level& L = levels_[level];
        dim3 blk(32, 8);
        if (op == 0) {
            if (level == nl - 1) throw Invalid{"coarsest level has no matrix-free operator exposed"};
            ApplyArgs<T> a = apply_args();
            a.x = in; a.y = out;
            launch_apply(L.g, level > 0, EP_PLAIN, a);
        } else if (op == 1) {
            if (level + 1 >= nl) throw Invalid{"no coarser level"};
            Level& C = levels_[level + 1];
            TM_CUDA(cudaMemsetAsync(out, 0, L.nu * sizeof(T), stream_));
            dim3 grd(ceil_div(L.g.Lx, 32), ceil_div(L.g.Ly, 8 * kProlongRows));
            mg_prolong_add_kernel<T><<<grd, blk, 0, stream_>>>(L.g, C.g, tr_tab_, in, out);
            TM_CHECK_LAUNCH();
        } else if (op == 2) {
            if (level + 1 >= nl) throw Invalid{"no coarser level"};
            Level& C = levels_[level + 1];
            dim3 grd(ceil_div(C.g.Lx, 32), ceil_div(C.g.Ly, 8));
            mg_restrict_kernel<T><<<grd, blk, 0, stream_>>>(L.g, C.g, tr_tab_, in, out);
            TM_CHECK_LAUNCH();
        } else if (op == 8) {  // the shared-memory tiled prolongation, whatever the level size (out += P in)
            if (level + 1 >= nl) throw Invalid{"no coarser level"};
            Level& C = levels_[level + 1];
            TM_CUDA(cudaMemsetAsync(out, 0, L.nu * sizeof(T), stream_));
            mg_prolong_tiled_kernel<T><<<dim3(ceil_div(L.g.Lx, kPtTI), ceil_div(L.g.Ly, kPtTJ)), kPtThreads, 0, stream_>>>(
                L.g, C.g, in, out);
            TM_CHECK_LAUNCH();
        } else if (op == 7) {  // the shared-memory tiled restriction, whatever the level size
            if (level + 1 >= nl) throw Invalid{"no coarser level"};
            Level& C = levels_[level + 1];
            mg_restrict_tiled_kernel<T><<<dim3(ceil_div(C.g.Lx, kRtTI), ceil_div(C.g.Ly, kRtTJ)), kRtThreads, 0, stream_>>>(
                L.g, C.g, in, out);
            TM_CHECK_LAUNCH();
        } else if (op == 3) {
            s_r_.ensure(nu_);
            TM_CUDA(cudaMemcpyAsync(s_r_.p, in, nu_ * sizeof(T), cudaMemcpyDeviceToDevice, stream_));
            const T* z = vcycle(s_r_.p);
            TM_CUDA(cudaMemcpyAsync(out, z, nu_ * sizeof(T), cudaMemcpyDeviceToDevice, stream_));
        } else if (op == 4) {
            if (level != nl - 1) throw Invalid{"op 4 is the coarsest-level direct solve"};
            mg_coarse_apply_inverse_kernel<T><<<1, 192, 0, stream_>>>((int)L.nu, coarse_Ainv_.p, (const T*)in, (T*)out);
            TM_CHECK_LAUNCH();
        } else if (op == 5) {
            if (level == nl - 1) throw Invalid{"coarsest level has no diagonal"};
            TM_CUDA(cudaMemcpyAsync(out, L.dinv.p, L.nu * sizeof(T), cudaMemcpyDeviceToDevice, stream_));
        } else if (op == 6) {
            // study: out[0] = mean ms of one V-cycle (graph replay when enabled, as in a solve),
            // out[1] = mean ms of one hierarchy set-up
            s_r_.ensure(nu_);
            TM_CUDA(cudaMemcpyAsync(s_r_.p, in, nu_ * sizeof(T), cudaMemcpyDeviceToDevice, stream_));
            cudaEvent_t e0, e1, e2;
            TM_CUDA(cudaEventCreate(&e0)); TM_CUDA(cudaEventCreate(&e1)); TM_CUDA(cudaEventCreate(&e2));
            for (int k = 0; k < 3; ++k) vcycle(s_r_.p);
            const int reps = 40, sreps = 5;
            TM_CUDA(cudaEventRecord(e0, stream_));
            for (int k = 0; k < reps; ++k) vcycle(s_r_.p);
            TM_CUDA(cudaEventRecord(e1, stream_));
            for (int k = 0; k < sreps; ++k) setup_hierarchy((T*)xi);
            TM_CUDA(cudaEventRecord(e2, stream_));
            TM_CUDA(cudaStreamSynchronize(stream_));
            float ms_v = 0.f, ms_s = 0.f;
            TM_CUDA(cudaEventElapsedTime(&ms_v, e0, e1));
            TM_CUDA(cudaEventElapsedTime(&ms_s, e1, e2));
            const T res[2] = {(T)(ms_v / reps), (T)(ms_s / sreps)};
            TM_CUDA(cudaMemcpyAsync(out, res, sizeof(res), cudaMemcpyHostToDevice, stream_));
            TM_CUDA(cudaStreamSynchronize(stream_));
            cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
        } else {
            throw Invalid{"unknown mg_debug op"};
        }
        TM_CUDA(cudaStreamSynchronize(stream_));
    }

    int mg_level_info(int level, int* info) override {
        if (level < 0 || level >= nlevels_) return nlevels_;
        const LevelGeom<T> g = level_geom(level, level >= dist_levels_);
        info[0] = g.nx; info[1] = g.nyg; info[2] = g.dl; info[3] = g.dr; info[4] = g.db; info[5] = g.dt;
        return nlevels_;
    }

   private:
    // ------------------------------------------------------------------ level planning
    // Global level sizes (ceil halving), Dirichlet thresholds, the number of sharded levels and
    // the partition of the level-0 cell rows (aligned so that strips nest across levels).
    void plan_levels(int dist_levels_hint) {
        lv_nx_.clear(); lv_ny_.clear(); lv_dr_.clear(); lv_dt_.clear();
        int nx = nx_, ny = nyg_;
        int dr = (cfg_.fixed_sides & TM_SIDE_RIGHT) ? 2 * nx_ : INT_MAX;
        int dt = (cfg_.fixed_sides & TM_SIDE_TOP) ? 2 * nyg_ : INT_MAX;
        while (true) {
            lv_nx_.push_back(nx); lv_ny_.push_back(ny); lv_dr_.push_back(dr); lv_dt_.push_back(dt);
            if (std::max(nx, ny) <= coarse_cells_) break;
            const int nxc = (nx + 1) / 2, nyc = (ny + 1) / 2;
            if (nxc == nx && nyc == ny) break;  // 1x1: cannot coarsen
            // far sides: last coarse column/row whose functions vanish on the fine Dirichlet nodes
            dr = dr == INT_MAX ? INT_MAX : 2 * (dr / 4);
            dt = dt == INT_MAX ? INT_MAX : 2 * (dt / 4);
            nx = nxc; ny = nyc;
        }
        nlevels_ = (int)lv_nx_.size();
        if (nlevels_ >= 2 &&
            2 * (size_t)(2 * lv_nx_.back() + 1) * (2 * lv_ny_.back() + 1) > (size_t)kCoarseMaxDofs)
            throw Invalid{"coarsest multigrid level too large"};
        if (nlevels_ > 30) throw Invalid{"too many multigrid levels"};

        // sharded levels: 0 .. dist_levels_-1
        dist_levels_ = 0;
        if (nranks_ > 1) {
            int ld = dist_levels_hint;
            if (ld <= 0) {
                ld = 1;
                while (ld < nlevels_ - 1 && (size_t)lv_nx_[ld] * lv_ny_[ld] > (size_t)262144 &&
                       (nyg_ >> (ld + 1)) >= 4 * nranks_)
                    ++ld;
            }
            dist_levels_ = std::max(1, std::min(ld, std::max(1, nlevels_ - 1)));
        }
        const int align = 1 << dist_levels_;
        starts_.assign(nranks_ + 1, nyg_);
        if (nranks_ == 1) {
            starts_[0] = 0;
        } else {
            const int per = ceil_div(ceil_div(nyg_, nranks_), align) * align;
            for (int r = 0; r < nranks_; ++r) starts_[r] = (int)std::min((long)r * per, (long)nyg_);
            if (starts_[nranks_ - 1] >= nyg_)
                throw Invalid{"mesh has too few cell rows for " + std::to_string(nranks_) + " ranks with " +
                              std::to_string(dist_levels_) + " sharded multigrid levels"};
        }
        ranges_.clear();
        for (int l = 0; l < nlevels_; ++l) ranges_.push_back(row_range(l, rank_, l >= dist_levels_));
    }

    void derive_local_sizes() {
        const RowRange& r0 = ranges_[0];
        ny_ = r0.ny();
        n1_ = (size_t)(nx_ + 1) * (ny_ + 1);
        n2_ = (size_t)(2 * nx_ + 1) * (2 * ny_ + 1);
        nu_ = 2 * n2_;
        p2_off_ = (size_t)r0.own_j0() * (2 * nx_ + 1) * 2;
        p2_cnt_ = (size_t)(r0.own_j1() - r0.own_j0()) * (2 * nx_ + 1) * 2;
        const int oi0 = r0.c0 - r0.cl0, oi1 = r0.c1 - r0.cl0 + (r0.last ? 1 : 0);
        p1_off_ = (size_t)oi0 * (nx_ + 1);
        p1_cnt_ = (size_t)(oi1 - oi0) * (nx_ + 1);
        // P1 element matrices (reference: FEM_src/filter.py:27-33; exact for P1)
        p1_.nx = nx_; p1_.ny = ny_; p1_.hx = hx_; p1_.hy = hy_;
        p1_.iy_off = r0.cl0; p1_.ny_global = nyg_; p1_.own_iy0 = oi0; p1_.own_iy1 = oi1;
        const double area = 0.5 * hx_ * hy_;
        const double gA[3][2] = {{-1 / hx_, 0}, {1 / hx_, -1 / hy_}, {0, 1 / hy_}};
        const double gB[3][2] = {{0, -1 / hy_}, {-1 / hx_, 1 / hy_}, {1 / hx_, 0}};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                p1_.Ke[0][i][j] = area * (gA[i][0] * gA[j][0] + gA[i][1] * gA[j][1]);
                p1_.Ke[1][i][j] = area * (gB[i][0] * gB[j][0] + gB[i][1] * gB[j][1]);
                p1_.Me[i][j] = area / 12.0 * (i == j ? 2.0 : 1.0);
            }
        p1_fill_interior_stencil(p1_);
        g0_ = level_geom(0, false);
    }

    RowRange row_range(int l, int rank, bool replicated) const {
        RowRange r;
        r.nyg = lv_ny_[l];
        if (replicated || nranks_ == 1) {
            r.c0 = 0; r.c1 = r.nyg; r.cl0 = 0; r.cl1 = r.nyg; r.last = true;
            return r;
        }
        r.last = rank == nranks_ - 1;
        r.c0 = starts_[rank] >> l;
        r.c1 = r.last ? r.nyg : (starts_[rank + 1] >> l);
        r.cl0 = std::max(0, r.c0 - 2);
        r.cl1 = std::min(r.nyg, r.c1 + 1);
        return r;
    }

    // rows of the FULL level l that `rank` owns (used to gather the first replicated level)
    void piece_rows(int l, int rank, int& j0, int& j1) const {
        const bool last = rank == nranks_ - 1;
        j0 = 2 * (starts_[rank] >> l);
        j1 = last ? 2 * lv_ny_[l] + 1 : 2 * (starts_[rank + 1] >> l);
    }

    LevelGeom<T> level_geom(int l, bool replicated) const {
        const RowRange r = row_range(l, rank_, replicated);
        LevelGeom<T> g;
        g.nx = lv_nx_[l]; g.ny = r.ny(); g.Lx = 2 * g.nx + 1; g.Ly = 2 * g.ny + 1;
        g.dl = (cfg_.fixed_sides & TM_SIDE_LEFT) ? 0 : -1;
        g.db = (cfg_.fixed_sides & TM_SIDE_BOTTOM) ? 0 : -1;
        g.dr = lv_dr_[l]; g.dt = lv_dt_[l];
        g.j_off = 2 * r.cl0; g.own_j0 = r.own_j0(); g.own_j1 = r.own_j1(); g.nyg = r.nyg;
        g.xi = nullptr; g.W = nullptr;
        g.simp_min = (T)cfg_.simp_min;
        g.mat = make_material<T>(cfg_.lame_lambda, cfg_.lame_mu, hx_, hy_);
        return g;
    }

    // ------------------------------------------------------------------ measurement ledger
    double sz(size_t n) const { return (double)n * sizeof(T); }
    void acct(int cat, double bytes) {
        led_.bytes[cat] += bytes;
        led_.launches[cat] += 1.0;
        if (profile_ == 3 && !capturing_) mark(cat);
    }
    void mark(int cat) {
        cudaEvent_t e = nullptr;
        if (ev_free_.empty()) {
            TM_CUDA(cudaEventCreate(&e));
        } else {
            e = ev_free_.back();
            ev_free_.pop_back();
        }
        TM_CUDA(cudaEventRecord(e, stream_));
        ev_marks_.push_back({cat, e});
    }
    // capture scope: launches recorded into a graph are filed when the graph is replayed
    struct CaptureLedger {
        Ledger saved;
        Ledger& live;
        bool& flag;
        CaptureLedger(Ledger& l, bool& f) : saved(l), live(l), flag(f) {
            live.clear();
            flag = true;
        }
        Ledger finish() {
            Ledger captured = live;
            live = saved;
            flag = false;
            return captured;
        }
    };

    // ------------------------------------------------------------------ helpers
    int grid1d(size_t n) const {
        const size_t want = (n + kVecThreads - 1) / kVecThreads;
        return (int)std::max<size_t>(1, std::min<size_t>(want, (size_t)num_sms_ * 8));
    }
    dim3 grid2d_p1() const { return dim3(ceil_div(nx_ + 1, 32), ceil_div(ny_ + 1, 8)); }

    void read_scalars() {
        TM_CUDA(cudaMemcpyAsync(h_sc_, sc_, sizeof(double) * SC_COUNT, cudaMemcpyDeviceToHost, stream_));
        unsigned int* err = reinterpret_cast<unsigned int*>(h_sc_ + 120);
        *err = 0;
        if (p2p_) {  // a timed-out peer poll surfaces with the same synchronisation
            const P2PHeader* h = reinterpret_cast<const P2PHeader*>(p2p_->win[rank_]);
            TM_CUDA(cudaMemcpyAsync(err, &h->error, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream_));
        }
        TM_CUDA(cudaStreamSynchronize(stream_));
        if (*err)
            throw Invalid{"peer-memory exchange timed out waiting for another rank (code " +
                          std::to_string(*err) + ")"};
    }

    void nccl_check(int rc, const char* what) {
        if (rc != 0) throw Invalid{std::string(what) + ": " + nccl().GetErrorString(rc)};
    }
    void need_comm() {
        if (!comm_) throw Invalid{"sharded engine used before tm_comm_init"};
    }

    // in-place sum over ranks of n device doubles
    void sum_ranks(double* p, int n) {
        if (nranks_ == 1) return;
        need_comm();
        if (p2p_ && n <= P2P_RED_MAX) {
            P2PReduceArgs a;
            for (int q = 0; q < P2P_MAX_RANKS; ++q) a.win[q] = q < nranks_ ? p2p_->win[q] : nullptr;
            a.rank = rank_;
            a.nranks = nranks_;
            a.p = p;
            a.n = n;
            p2p_allreduce_kernel<<<1, 64, 0, stream_>>>(a);
            TM_CHECK_LAUNCH();
            acct(LC_ALLREDUCE, 8.0 * n * nranks_);
            return;
        }
        nccl_check(nccl().AllReduce(p, p, (size_t)n, kNcclFloat64, kNcclSum, comm_, stream_), "ncclAllReduce");
        acct(LC_ALLREDUCE, 8.0 * n * nranks_);
    }

    // halo exchange of a row-major local array: the `up` rows below my top edge go to rank+1,
    // which stores them in its first rows; `down` rows from my first owned row go to rank-1
    void exchange_rows(T* v, size_t row_elems, int own0, int own1, int up, int down, bool last) {
        if (nranks_ == 1) return;
        need_comm();
        if (p2p_ && (size_t)std::max(up, down) * row_elems * sizeof(T) <= p2p_->slot_bytes) {
            P2PHaloArgs a;
            const size_t rb = row_elems * sizeof(T);
            a.self = p2p_->win[rank_];
            a.below = rank_ > 0 ? p2p_->win[rank_ - 1] : nullptr;
            a.above = !last ? p2p_->win[rank_ + 1] : nullptr;
            a.slot_bytes = p2p_->slot_bytes;
            a.v = reinterpret_cast<char*>(v);
            a.up_src = (size_t)(own1 - up) * rb;
            a.up_bytes = (size_t)up * rb;
            a.down_src = (size_t)own0 * rb;
            a.down_bytes = (size_t)down * rb;
            a.from_below_dst = (size_t)(own0 - up) * rb;
            a.from_above_dst = (size_t)own1 * rb;
            const size_t big = std::max(a.up_bytes, a.down_bytes);
            const uintptr_t align = reinterpret_cast<uintptr_t>(v) | a.up_src | a.up_bytes | a.down_src |
                                    a.down_bytes | a.from_below_dst | a.from_above_dst;
            const int granule = (align % 16 == 0) ? 16 : (align % 8 == 0) ? 8 : 4;
            const int blocks = (int)std::min<size_t>((size_t)num_sms_, (big / granule + 1023) / 1024 + 1);
            if (granule == 16) p2p_halo_kernel<uint4><<<blocks, 256, 0, stream_>>>(a);
            else if (granule == 8) p2p_halo_kernel<uint2><<<blocks, 256, 0, stream_>>>(a);
            else p2p_halo_kernel<unsigned int><<<blocks, 256, 0, stream_>>>(a);
            TM_CHECK_LAUNCH();
            // rows pushed to the neighbours + rows unpacked from the own mailboxes
            acct(LC_HALO, 2.0 * ((!last ? 1 : 0) + (rank_ > 0 ? 1 : 0)) * (double)(up + down) * row_elems * sizeof(T));
            return;
        }
        const int dtype = sizeof(T) == 8 ? kNcclFloat64 : kNcclFloat32;
        nccl_check(nccl().GroupStart(), "ncclGroupStart");
        if (!last) {
            nccl().Send(v + (size_t)(own1 - up) * row_elems, (size_t)up * row_elems, dtype, rank_ + 1, comm_, stream_);
            nccl().Recv(v + (size_t)own1 * row_elems, (size_t)down * row_elems, dtype, rank_ + 1, comm_, stream_);
        }
        if (rank_ > 0) {
            nccl().Send(v + (size_t)own0 * row_elems, (size_t)down * row_elems, dtype, rank_ - 1, comm_, stream_);
            nccl().Recv(v + (size_t)(own0 - up) * row_elems, (size_t)up * row_elems, dtype, rank_ - 1, comm_, stream_);
        }
        nccl_check(nccl().GroupEnd(), "ncclGroupEnd");
        acct(LC_HALO, 2.0 * ((!last ? 1 : 0) + (rank_ > 0 ? 1 : 0)) * (double)(up + down) * row_elems * sizeof(T));
    }
    // lattice vector of sharded level l: 4 rows up, 3 rows down
    void exchange_p2(int l, T* v) {
        if (nranks_ == 1 || l >= dist_levels_) return;
        const RowRange& r = ranges_[l];
        exchange_rows(v, (size_t)(2 * lv_nx_[l] + 1) * 2, r.own_j0(), 2 * (r.c1 - r.cl0), 4, 3, r.last);
    }
    // P1 field (level 0): 2 vertex rows each way
    void exchange_p1(T* v) {
        if (nranks_ == 1) return;
        const RowRange& r = ranges_[0];
        exchange_rows(v, (size_t)nx_ + 1, r.c0 - r.cl0, r.c1 - r.cl0, 2, 2, r.last);
    }
    // stored moments of sharded level l >= 1: 12 planes, 2 cell rows up, 1 down
    void exchange_w(int l, T* W) {
        if (nranks_ == 1 || l >= dist_levels_) return;
        const RowRange& r = ranges_[l];
        const size_t plane = (size_t)lv_nx_[l] * r.ny();
        for (int k = 0; k < 12; ++k)
            exchange_rows(W + k * plane, (size_t)lv_nx_[l], r.c0 - r.cl0, r.c1 - r.cl0, 2, 1, r.last);
    }
    // every rank contributes its rows of the FULL array (row_elems per row) of level l
    void gather_rows(int l, T* full, size_t row_elems, bool cell_rows) {
        if (nranks_ == 1) return;
        need_comm();
        const int dtype = sizeof(T) == 8 ? kNcclFloat64 : kNcclFloat32;
        nccl_check(nccl().GroupStart(), "ncclGroupStart");
        for (int q = 0; q < nranks_; ++q) {
            int j0, j1;
            if (cell_rows) {
                j0 = starts_[q] >> l;
                j1 = q == nranks_ - 1 ? lv_ny_[l] : (starts_[q + 1] >> l);
            } else {
                piece_rows(l, q, j0, j1);
            }
            if (j1 <= j0) continue;
            T* p = full + (size_t)j0 * row_elems;
            nccl().Broadcast(p, p, (size_t)(j1 - j0) * row_elems, dtype, q, comm_, stream_);
        }
        nccl_check(nccl().GroupEnd(), "ncclGroupEnd");
        acct(LC_GATHER, sz((size_t)(cell_rows ? lv_ny_[l] : 2 * lv_ny_[l] + 1) * row_elems));
    }

    ApplyArgs<T> apply_args() const {
        ApplyArgs<T> a;
        std::memset(&a, 0, sizeof(a));
        a.rs = rs_;
        a.store_d = 1;
        return a;
    }

    // Launch of a kernel that starts with pdl_prologue(): inside a V-cycle (pdl_active_) it carries
    // the programmatic-stream-serialization attribute, so that under stream capture the graph gets
    // programmatic edges and each kernel's launch overlaps the tail of its predecessor.
    template <class... P, class... A>
    void launch_chain(void (*kernel)(P...), dim3 grd, dim3 blk, A&&... args) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = grd;
        cfg.blockDim = blk;
        cfg.dynamicSmemBytes = 0;
        cfg.stream = stream_;
        cudaLaunchAttribute at{};
        at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &at;
        cfg.numAttrs = pdl_active_ ? 1 : 0;
        TM_CUDA(cudaLaunchKernelEx(&cfg, kernel, std::forward<A>(args)...));
    }

    void launch_apply(const LevelGeom<T>& g, bool stored, int ep, ApplyArgs<T> a) {
        stored = stored || g.W != nullptr;  // level 0 with a general SIMP exponent carries stored moments
        const int ncg = ceil_div(g.nx + 1, 31);
        const int bx = ceil_div(ncg, kApplyWarps);
        // A strip re-evaluates the cell row below it ((H+1)/H flops) and marches H rows
        // serially: large meshes get H up to 64, small (latency-bound) levels H down to 1 so
        // that there are always ~4 blocks per SM to overlap the march latencies.
        int strips = ceil_div((long)blocks_per_sm_target_ * num_sms_, bx);
        // The density-driven fine-level kernel holds apply_minb_ blocks per SM (250 registers): size
        // its grid to ONE full wave when the mesh allows, never slightly more (A/B on N=512: 288
        // blocks of 16 rows against 576 of 8: 3 % faster, less pre-roll; the stored-moment levels
        // are better off with ~4 blocks per SM)
        if (!stored) strips = std::max(1, (int)((long)apply_minb_ * num_sms_ / bx));
        strips = std::min(strips, g.ny);
        strips = std::max(strips, ceil_div(g.ny, 64));
        a.rows_per_strip = std::max(min_rows_per_strip_, ceil_div(g.ny, strips));
        strips = ceil_div(g.ny, a.rows_per_strip);
        if (stored && wave_aware_) {
            // Stored-moment levels: pick the strip height whose grid fills whole waves of resident blocks
            // (bridge N=2048, level 2: 600 blocks on 444 resident slots = 1.35 waves ran at 3.6 TB/s, level 1
            // at 1.8 waves at 4.8 TB/s, and 448 blocks = 1.01 waves costs two; profiles/r2b), weighed against
            // the pre-roll row every strip re-evaluates.
            if (stored_cap_ == 0) {
                int per_sm = 0;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                        &per_sm, elast_apply_kernel<T, true, EP_CHEB, 3, false>, kApplyWarps * 32, 0) != cudaSuccess ||
                    per_sm < 1) {
                    cudaGetLastError();
                    per_sm = 3;
                }
                stored_cap_ = (long)per_sm * num_sms_;
            }
            // time model: whole waves of resident blocks, each marching h + 1 cell rows (a partly filled last
            // wave costs a full one; a level below one wave is latency bound: the shortest march wins)
            long best_t = -1;
            int best_h = a.rows_per_strip;
            for (int h = 64; h >= min_rows_per_strip_; --h) {
                const long nb = (long)bx * ceil_div(g.ny, h);
                const long t = (long)ceil_div(nb, stored_cap_) * (h + 1);
                if (best_t < 0 || t < best_t) {  // ties: the taller strip (fewer redundant rows)
                    best_t = t;
                    best_h = h;
                }
            }
            a.rows_per_strip = best_h;
            strips = ceil_div(g.ny, a.rows_per_strip);
        }
        dim3 grd(bx, strips), blk(kApplyWarps * 32);
        if ((long)bx * strips > rs_.capacity) throw Invalid{"reduction scratch too small"};
        const bool fine = g.nx == nx_ && g.nyg == nyg_;
        int level = 0;
        while (level + 1 < nlevels_ && !(lv_nx_[level] == g.nx && lv_ny_[level] == g.nyg)) ++level;
        std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
        const bool timed = profile_ == 2 || (profile_ == 1 && fine);
        // statistics file the fused variants under their base epilogue
        const int ep_slot = ep == EP_RESID0 ? (int)EP_RESID : ((ep == EP_CHEBDOT || ep == EP_CHEB0) ? (int)EP_CHEB : ep);
        if (timed) {
            if (prof_free_.empty()) {
                TM_CUDA(cudaEventCreate(&ev.first));
                TM_CUDA(cudaEventCreate(&ev.second));
            } else {
                ev = prof_free_.back();
                prof_free_.pop_back();
            }
            TM_CUDA(cudaEventRecord(ev.first, stream_));
        }
#define TM_LAUNCH_APPLY(ST, EPV, MB, PFV) launch_chain(elast_apply_kernel<T, ST, EPV, MB, PFV>, grd, blk, g, a)
#define TM_LAUNCH_APPLY_EP(ST, MB, PFV)                               \
    switch (ep) {                                                     \
        case EP_PLAIN: TM_LAUNCH_APPLY(ST, EP_PLAIN, MB, PFV); break;   \
        case EP_DOT: TM_LAUNCH_APPLY(ST, EP_DOT, MB, PFV); break;       \
        case EP_RESID: TM_LAUNCH_APPLY(ST, EP_RESID, MB, PFV); break;   \
        case EP_RESID0: TM_LAUNCH_APPLY(ST, EP_RESID0, MB, PFV); break; \
        case EP_CHEBDOT: TM_LAUNCH_APPLY(ST, EP_CHEBDOT, MB, PFV); break; \
        case EP_CHEB0: TM_LAUNCH_APPLY(ST, EP_CHEB0, MB, PFV); break;   \
        default: TM_LAUNCH_APPLY(ST, EP_CHEB, MB, PFV); break;          \
    }
        if (stored) {
            TM_LAUNCH_APPLY_EP(true, 3, false)
        } else if (apply_minb_ >= 4) {
            TM_LAUNCH_APPLY_EP(false, 4, true)
        } else {
            TM_LAUNCH_APPLY_EP(false, 2, true)
        }
#undef TM_LAUNCH_APPLY_EP
#undef TM_LAUNCH_APPLY
        TM_CHECK_LAUNCH();
        {   // algorithmic bytes: every lattice vector the epilogue touches once + the coefficient source
            const double nuv = 2.0 * g.Lx * g.Ly;
            double vecs = 2.0;  // EP_PLAIN, EP_DOT: read x, write y
            if (ep == EP_RESID) vecs = 3.0;  // + b
            else if (ep == EP_RESID0) vecs = 4.0;  // read b, D^-1; write x, r
            else if (ep == EP_CHEB0) vecs = 3.0 + (a.store_d ? 1.0 : 0.0);  // read b, D^-1; write y (, d)
            else if (ep_is_cheb(ep)) vecs = 4.0 + (a.c1 != T(0) ? 1.0 : 0.0) + (a.store_d ? 1.0 : 0.0);
            const double coef = stored ? 12.0 * g.nx * g.ny : (double)(g.nx + 1) * (g.ny + 1);
            const int fine_cat = ep == EP_CHEB0 ? (int)LC_FINE_CHEB : LC_FINE_PLAIN + ep;
            acct(fine ? fine_cat : LC_COARSE1 + std::min(level, 14) - 1, (vecs * nuv + coef) * sizeof(T));
        }
        if (timed) {
            TM_CUDA(cudaEventRecord(ev.second, stream_));
            prof_pending_.push_back({4 * level + ep_slot, ev});
        }
        if (fine) {
            ++stats_fine_applies_;
            ++fine_ep_count_[ep_slot];
        }
    }

    void launch_diag(const LevelGeom<T>& g, bool stored, T* dinv) {
        stored = stored || g.W != nullptr;
        dim3 blk(32, 8), grd(ceil_div(g.Lx, 32), ceil_div(g.Ly, 8));
        if (stored)
            elast_diag_kernel<T, true><<<grd, blk, 0, stream_>>>(g, diag_tab_, dinv);
        else
            elast_diag_kernel<T, false><<<grd, blk, 0, stream_>>>(g, diag_tab_, dinv);
        TM_CHECK_LAUNCH();
        acct(LC_SETUP_DIAG, (2.0 * g.Lx * g.Ly + (stored ? 12.0 * g.nx * g.ny : (double)(g.nx + 1) * (g.ny + 1))) * sizeof(T));
    }

    void p1_apply(double alpha, double beta, const T* x, T* y, double* dot_out) {
        dim3 grd = grid2d_p1();
        if ((long)grd.x * grd.y > rs_.capacity) throw Invalid{"reduction scratch too small"};
        if (dot_out)
            p1_apply_kernel<T, true><<<grd, dim3(32, 8), 0, stream_>>>(p1_, alpha, beta, x, y, rs_, dot_out);
        else
            p1_apply_kernel<T, false><<<grd, dim3(32, 8), 0, stream_>>>(p1_, alpha, beta, x, y, rs_, nullptr);
        TM_CHECK_LAUNCH();
        acct(LC_FILTER, 2 * sz(n1_));
    }

    // ------------------------------------------------------------------ PCG
    // Solves A x = b given r = b - A x on entry.  All vectors are local arrays; the flat vector
    // kernels touch the owned range [off, off+n) only.  precond(r) returns z = M^-1 r, or
    // nullptr when `jacobi` (then z = dinv r is fused into the vector kernels).
    template <class ApplyDot, class Precond>
    SolveStats pcg(size_t off, size_t n, const T* b, T* x, T* r, T* p, T* Ap, const T* dinv,
                   ApplyDot apply_dot, Precond precond, bool jacobi, double rtol, int maxit, int check,
                   bool filter_solve = false, int check_from = 0, double floor_factor = 0.0) {
        SolveStats st;
        const int g1 = grid1d(n);
        const T* dv = dinv ? dinv + off : nullptr;
        // ledger categories of the vector kernels (the filter's solves are filed as a whole)
        const int lc_upd = filter_solve ? LC_FILTER : LC_PCG_UPDATE, lc_dir = filter_solve ? LC_FILTER : LC_PCG_DIRECTION,
                  lc_red = filter_solve ? LC_FILTER : LC_REDUCTIONS;
        dot_kernel<T><<<g1, kVecThreads, 0, stream_>>>(n, b + off, b + off, rs_, sc_ + SC_BB);
        TM_CHECK_LAUNCH();
        acct(lc_red, sz(n));
        sum_ranks(sc_ + SC_BB, 1);
        int cur = 0;
        if (jacobi) {
            pcg_start_kernel<T, true><<<g1, kVecThreads, 0, stream_>>>(n, sc_, cur, p + off, r + off, dv, rs_);
        } else {
            const T* z = precond(r);
            pcg_start_kernel<T, false><<<g1, kVecThreads, 0, stream_>>>(n, sc_, cur, p + off, r + off,
                                                                       z + off, rs_);
        }
        TM_CHECK_LAUNCH();
        acct(lc_red, 3 * sz(n));
        sum_ranks(sc_ + SC_RR, 1);
        sum_ranks(sc_ + cur, 1);
        read_scalars();
        const double bb = h_sc_[SC_BB];
        if (!(bb > 0.0)) {
            TM_CUDA(cudaMemsetAsync(x + off, 0, n * sizeof(T), stream_));
            st.converged = true;
            return st;
        }
        if (floor_factor > 0.0) {
            // attainable accuracy: sc_[SC_FLOOR] = ||A w||^2 for w = the half-ulp perturbation of the initial
            // guess.  The recursive residual of CG keeps falling below that level, the true one b - A x cannot;
            // iterating further only spends time (Greenbaum 1997; measured in DESIGN.md section 5)
            st.floor = std::sqrt(std::max(h_sc_[SC_FLOOR], 0.0) / bb);
            if (st.floor == st.floor) rtol = std::max(rtol, floor_factor * st.floor);
        }
        st.rtol_used = rtol;
        st.relres = std::sqrt(h_sc_[SC_RR] / bb);
        if (st.relres <= rtol) {
            st.converged = true;
            return st;
        }
        for (int k = 0; k < maxit; ++k) {
            apply_dot(p, Ap);
            if (jacobi) {
                pcg_update_kernel<T, true><<<g1, kVecThreads, 0, stream_>>>(
                    n, sc_, cur, cur ^ 1, x + off, r + off, p + off, Ap + off, dv, rs_);
                TM_CHECK_LAUNCH();
                acct(lc_upd, 7 * sz(n));
                sum_ranks(sc_ + SC_RR, 1);
                sum_ranks(sc_ + (cur ^ 1), 1);
            } else {
                pcg_update_kernel<T, false><<<g1, kVecThreads, 0, stream_>>>(
                    n, sc_, cur, cur ^ 1, x + off, r + off, p + off, Ap + off, nullptr, rs_);
                TM_CHECK_LAUNCH();
                acct(lc_upd, 6 * sz(n));
                sum_ranks(sc_ + SC_RR, 1);
            }
            st.iters = k + 1;
            // check_from: no residual read-back (a host synchronisation) before that iteration, apart from a
            // sparse safety net -- the caller expects about as many iterations as its previous solve took
            const bool do_check = (k + 1 >= check_from && (k + 1) % check == 0) || (k + 1) % 16 == 0 || (k + 1 == maxit);
            if (do_check) {
                read_scalars();
                st.relres = std::sqrt(h_sc_[SC_RR] / bb);
                if (!(st.relres == st.relres)) break;  // NaN: give up
                if (st.relres <= rtol) {
                    st.converged = true;
                    break;
                }
            }
            if (jacobi) {
                pcg_direction_kernel<T, true><<<g1, kVecThreads, 0, stream_>>>(n, sc_, cur, cur ^ 1, -1, p + off,
                                                                              r + off, dv);
            } else {
                vcycle_rz_ = false;
                const T* z = precond(r);
                if (vcycle_rz_) {  // r . z came out of the V-cycle's last smoothing step
                    sum_ranks(sc_ + SC_RZV, 1);
                    pcg_direction_kernel<T, false><<<g1, kVecThreads, 0, stream_>>>(n, sc_, cur, SC_RZV, cur ^ 1,
                                                                                   p + off, r + off, z + off);
                } else {
                    dot_kernel<T><<<g1, kVecThreads, 0, stream_>>>(n, r + off, z + off, rs_, sc_ + (cur ^ 1));
                    TM_CHECK_LAUNCH();
                    acct(lc_red, 2 * sz(n));
                    sum_ranks(sc_ + (cur ^ 1), 1);
                    pcg_direction_kernel<T, false><<<g1, kVecThreads, 0, stream_>>>(n, sc_, cur, cur ^ 1, -1,
                                                                                   p + off, r + off, z + off);
                }
            }
            TM_CHECK_LAUNCH();
            acct(lc_dir, (jacobi ? 4 : 3) * sz(n));
            cur ^= 1;
        }
        return st;
    }

    // ------------------------------------------------------------------ multigrid
    struct Level {
        LevelGeom<T> g;       // local geometry (own rows = this rank's)
        LevelGeom<T> gpiece;  // first replicated level only: full geometry, own rows = my piece
        size_t nu = 0, off = 0, cnt = 0;  // local size; owned flat range
        bool sharded = false;
        DevBuf<T> W, dinv, x, xalt, b, d, tmp, eig;
        DevBuf<T> S;  // tail levels only: assembled node stencils (tm_tail.cuh)
        bool eig_ready = false;
        double lmax = 0.0;
    };

    void build_levels() {
        if (setup_graph_exec_) {
            cudaGraphExecDestroy(setup_graph_exec_);
            setup_graph_exec_ = nullptr;
        }
        levels_.clear();
        levels_.resize(nlevels_);
        for (int l = 0; l < nlevels_; ++l) {
            Level& L = levels_[l];
            L.sharded = l < dist_levels_;
            L.g = level_geom(l, !L.sharded);
            L.gpiece = L.g;
            L.nu = 2 * (size_t)L.g.Lx * L.g.Ly;
            L.off = (size_t)L.g.own_j0 * L.g.Lx * 2;
            L.cnt = (size_t)(L.g.own_j1 - L.g.own_j0) * L.g.Lx * 2;
            if (nranks_ > 1 && l == dist_levels_) {
                int j0, j1;
                piece_rows(l, rank_, j0, j1);
                L.gpiece.own_j0 = j0;
                L.gpiece.own_j1 = j1;
            }
            const bool coarsest = (l + 1 == nlevels_);
            if (l > 0) {
                L.W.ensure(12 * (size_t)L.g.nx * L.g.ny);
                L.g.W = L.W.p;
                L.gpiece.W = L.W.p;
                L.b.ensure(L.nu);
            }
            L.x.ensure(L.nu);
            if (!coarsest) {
                L.dinv.ensure(L.nu); L.xalt.ensure(L.nu); L.d.ensure(L.nu);
                L.tmp.ensure(L.nu); L.eig.ensure(L.nu);
            }
        }
        const size_t nc = levels_.back().nu;
        coarse_Ainv_.ensure(nc * nc);
        coarse_inv_smem_ = (nc * nc + 2 * nc) * sizeof(double);
        TM_CUDA(cudaFuncSetAttribute(mg_coarse_inverse_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)coarse_inv_smem_));
        {   // stencil lists of the tiled restriction (tm_mg.cuh)
            const RestrictLists rl = make_restrict_lists(tr_tab_);
            TM_CUDA(cudaMemcpyToSymbol(c_restrict_lists, &rl, sizeof(rl)));
            const ProlongLists pl = make_prolong_lists(tr_tab_);
            TM_CUDA(cudaMemcpyToSymbol(c_prolong_lists, &pl, sizeof(pl)));
        }
        plan_tail();
    }

    static const void* filter_tb_entry(int rows) {
        switch (rows) {
            case 2: return (const void*)filter_cheb_tb_kernel<T, 2>;
            case 4: return (const void*)filter_cheb_tb_kernel<T, 4>;
            case 6: return (const void*)filter_cheb_tb_kernel<T, 6>;
            case 8: return (const void*)filter_cheb_tb_kernel<T, 8>;
            case 10: return (const void*)filter_cheb_tb_kernel<T, 10>;
            case 12: return (const void*)filter_cheb_tb_kernel<T, 12>;
            default: return (const void*)filter_cheb_tb_kernel<T, 14>;
        }
    }
    // Chebyshev coefficients (c1, c2) of iterations 0 .. len-1 for the current spectral bounds
    const double* filter_c12(int len) {
        if ((int)filter_c12_host_.size() < 2 * len || filter_c12_lmin_ != filter_lmin_ ||
            filter_c12_lmax_ != filter_lmax_) {
            len = std::max(len, 2048);
            filter_c12_host_.resize(2 * (size_t)len);
            const double theta = 0.5 * (filter_lmax_ + filter_lmin_), delta = 0.5 * (filter_lmax_ - filter_lmin_);
            const double sigma = theta / delta;
            double rho = 1.0 / sigma;
            filter_c12_host_[0] = 0.0;
            filter_c12_host_[1] = 1.0 / theta;
            for (int k = 1; k < len; ++k) {
                const double rho_new = 1.0 / (2.0 * sigma - rho);
                filter_c12_host_[2 * k] = rho_new * rho;
                filter_c12_host_[2 * k + 1] = 2.0 * rho_new / delta;
                rho = rho_new;
            }
            filter_c12_.ensure(2 * (size_t)len);
            TM_CUDA(cudaMemcpyAsync(filter_c12_.p, filter_c12_host_.data(), sizeof(double) * 2 * len,
                                    cudaMemcpyHostToDevice, stream_));
            TM_CUDA(cudaStreamSynchronize(stream_));
            filter_c12_lmin_ = filter_lmin_;
            filter_c12_lmax_ = filter_lmax_;
        }
        return filter_c12_.p;
    }

    // Tiling of the vertex grid for the temporally blocked Chebyshev filter (tm_filter_pcg.cuh):
    // at most one tile per SM; the two padded x arrays of the extended tile must fit shared
    // memory and a thread's column segment its register arrays.  Among those, the tiling with
    // the shortest per-thread march (the critical path of a step), then the smallest tile.
    bool plan_filter_tb(FilterTbArgs& a) {
        if (filter_tb_state_ < 0) return false;
        const int S = std::max(1, filter_tb_steps_);
        if (filter_tb_state_ == 0) {
            filter_tb_state_ = -1;
            int smem_max = 0;
            TM_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
            const long budget = ((long)smem_max - 4096) / (2 * (long)sizeof(double));  // doubles per array
            const int W1 = p1_.nx + 1, H1 = p1_.ny + 1;
            long best = -1;
            for (int tx = 1; tx <= num_sms_; ++tx)
                for (int ty = 1; tx * ty <= num_sms_; ++ty) {
                    const int tw = ceil_div(W1, tx), th = ceil_div(H1, ty);
                    if (ceil_div(W1, tw) != tx || ceil_div(H1, th) != ty) continue;  // no empty tiles
                    const int EW = tw + 2 * S, EH = th + 2 * S;
                    if (EW > kFilterTbThreads) continue;
                    const int nseg = kFilterTbThreads / EW;
                    const int rows = 2 * ceil_div(ceil_div(EH, nseg), 2);  // instantiated: 2, 4, .. 14
                    if (rows > kFilterTbMaxRows || (long)(EW + 2) * (nseg * rows + 2) > budget) continue;
                    const long cost = (long)rows * 1000000 + (long)EW * EH;
                    if (best < 0 || cost < best) {
                        best = cost;
                        filter_tb_tx_ = tx; filter_tb_ty_ = ty; filter_tb_tw_ = tw; filter_tb_th_ = th;
                        filter_tb_rows_ = rows;
                        filter_tb_smem_ = (size_t)(EW + 2) * (nseg * rows + 2) * 2 * sizeof(double);
                    }
                }
            if (best < 0) return false;
            const void* entry = filter_tb_entry(filter_tb_rows_);
            if (cudaFuncSetAttribute(entry, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)filter_tb_smem_) !=
                cudaSuccess) {
                cudaGetLastError();
                return false;
            }
            int per_sm = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, entry, kFilterTbThreads, filter_tb_smem_) !=
                    cudaSuccess || per_sm < 1) {
                cudaGetLastError();
                return false;
            }
            filter_tb_state_ = 1;
        }
        a.g = p1_;
        a.S = S;
        a.tiles_x = filter_tb_tx_; a.tiles_y = filter_tb_ty_; a.tw = filter_tb_tw_; a.th = filter_tb_th_;
        a.rows_per_thread = filter_tb_rows_;
        return true;
    }

    // The levels from tail_first_ down run as one cluster kernel per V-cycle (tm_tail.cuh):
    // replicated levels with stored moments whose lattice has at most tail_max_nodes_ nodes.
    void plan_tail() {
        tail_first_ = -1;
        const int degree = coarse_degree_ > 0 ? coarse_degree_ : cheb_degree_;
        if (tail_max_nodes_ <= 0 || degree > kTailMaxDegree) return;
        int first = -1;
        for (int l = std::max(1, nranks_ > 1 ? dist_levels_ : 1); l < nlevels_; ++l)
            if ((long)levels_[l].g.Lx * levels_[l].g.Ly <= (long)tail_max_nodes_) {
                first = l;
                break;
            }
        if (first < 0 || nlevels_ - first < 2) return;
        first = std::max(first, nlevels_ - kTailMaxLevels);
        int smem_max = 0;
        TM_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        const size_t arg_bytes = (sizeof(TailArgs<T>) + 15) / 16 * 16;
        for (int csize = tail_cluster_; csize >= 1; csize /= 2) {
            // stencil images of the smoothed tail levels for this cluster size
            std::memset(&tail_host_, 0, sizeof(tail_host_));
            int img_len = 0;
            for (int l = first; l + 1 < nlevels_; ++l) {
                TailLevel<T>& V = tail_host_.lv[l - first];
                V.g = levels_[l].g;
                img_len += tail_plan_level(V, csize, img_len);
            }
            const size_t smem = arg_bytes + (size_t)img_len * sizeof(T);
            if (smem > (size_t)smem_max) continue;
            if (cudaFuncSetAttribute(tail_vcycle_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem) != cudaSuccess ||
                (csize > 8 && cudaFuncSetAttribute(tail_vcycle_kernel<T>,
                                                   cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)) {
                cudaGetLastError();
                continue;
            }
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(csize);
            cfg.blockDim = dim3(kTailThreads);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = csize;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            int nclusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nclusters, tail_vcycle_kernel<T>, &cfg) != cudaSuccess ||
                nclusters < 1) {
                cudaGetLastError();
                continue;
            }
            tail_cluster_used_ = csize;
            tail_first_ = first;
            tail_smem_ = smem;
            tail_host_.img_len = img_len;
            tail_img_.ensure((size_t)csize * img_len);
            tail_dev_.ensure(sizeof(TailArgs<T>));
            return;
        }
    }

    // stencils of the tail levels (after the moments were coarsened) ...
    void assemble_tail() {
        if (tail_first_ < 0) return;
        for (int l = tail_first_; l + 1 < nlevels_; ++l) {
            TailLevel<T>& V = tail_host_.lv[l - tail_first_];
            V.g = levels_[l].g;  // W pointer is current
            tail_assemble_kernel<T><<<ceil_div(V.toff[4] * V.npt * 2, 128), 128, 0, stream_>>>(
                V, tail_cluster_used_, tail_host_.img_len, tail_img_.p);
            TM_CHECK_LAUNCH();
            acct(LC_SETUP_COARSEN, 12 * sz((size_t)V.g.nx * V.g.ny) + sz((size_t)tail_cluster_used_ * tail_host_.img_len));
        }
    }
    // ... and the argument block (needs the smoother bounds, i.e. the host copy of lambda_max)
    void upload_tail() {
        if (tail_first_ < 0) return;
        TailArgs<T>& A = tail_host_;
        const int nl = nlevels_;
        A.nt = nl - tail_first_;
        A.degree = coarse_degree_ > 0 ? coarse_degree_ : cheb_degree_;
        A.nc = (int)levels_[nl - 1].nu;
        A.dry = tail_dry_;
        A.img = tail_img_.p;
        A.Ainv = coarse_Ainv_.p;
        A.tab = tr_tab_;
        for (int t = 0; t < A.nt; ++t) {
            Level& L = levels_[tail_first_ + t];
            TailLevel<T>& V = A.lv[t];
            V.g = L.g;
            V.n = L.g.Lx * L.g.Ly;
            V.dinv = L.dinv.p;
            V.b = L.b.p;
            V.xa = L.x.p;
            V.xb = L.xalt.p;
            V.d = L.d.p;
            V.r = L.tmp.p;
            V.gamma = repeats(tail_first_ + t);
            if (t + 1 == A.nt) {  // coarsest level: only its lattice and the transfer split are used
                if (V.G == 0) tail_plan_level(V, tail_cluster_used_, 0);
                break;
            }
            const double hi = eig_safety_ * L.lmax, lo = hi / cheb_ratio_;
            const double theta = 0.5 * (hi + lo), delta = 0.5 * (hi - lo), sigma = theta / delta;
            double rho = 1.0 / sigma;
            V.c1[0] = T(0);
            V.c2[0] = (T)(1.0 / theta);
            for (int k = 1; k < A.degree; ++k) {
                const double rho_new = 1.0 / (2.0 * sigma - rho);
                V.c1[k] = (T)(rho_new * rho);
                V.c2[k] = (T)(2.0 * rho_new / delta);
                rho = rho_new;
            }
        }
        TM_CUDA(cudaMemcpyAsync(tail_dev_.p, &A, sizeof(A), cudaMemcpyHostToDevice, stream_));
        TM_CUDA(cudaStreamSynchronize(stream_));
    }
    // x of level tail_first_ after the kernel: the ping-pong buffer the last step wrote.  x0 != nullptr:
    // the cycle starts from that guess (one of the level's two iterate buffers) and returns in the same buffer
    T* launch_tail(T* x0 = nullptr) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(tail_cluster_used_);
        cfg.blockDim = dim3(kTailThreads);
        cfg.dynamicSmemBytes = tail_smem_;
        cfg.stream = stream_;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = tail_cluster_used_;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        Level& L = levels_[tail_first_];
        if (x0 && x0 != L.x.p && x0 != L.xalt.p) throw Invalid{"tail: the initial guess is not an iterate buffer"};
        const int warm = !x0 ? 0 : (x0 == L.x.p ? 1 : 2);
        TM_CUDA(cudaLaunchKernelEx(&cfg, tail_vcycle_kernel<T>,
                                   reinterpret_cast<const TailArgs<T>*>(tail_dev_.p), warm));
        TM_CHECK_LAUNCH();
        acct(LC_TAIL, 2 * sz(levels_[tail_first_].nu));
        if (x0) return x0;  // D pre- and D post-smoothing steps: an even number of buffer flips
        const int D = tail_host_.degree;
        const int flips = std::max(D - 2, 0) + D;
        return (flips & 1) ? L.xalt.p : L.x.p;
    }

    // Coefficients of every level, smoother bounds, coarsest factorisation, tail stencils.  The
    // ~150 small launches of the device part are captured into a CUDA graph once the warm-started
    // eigenvector iteration has its steady shape, and replayed while `xi` is the same buffer.
    void setup_hierarchy(T* xi) {
        const int nl = nlevels_;
        graph_dirty_ = true;  // coefficients / smoother bounds change: re-capture the V-cycle
        graph_sampled_ = false;
        bool steady = use_graph_ && nranks_ == 1 && profile_ < 2;
        for (int l = 0; steady && l + 1 < nl; ++l) steady = levels_[l].eig_ready;
        if (!steady) {
            setup_device_part(xi);
        } else if (setup_graph_exec_ && setup_graph_xi_ == xi) {
            TM_CUDA(cudaGraphLaunch(setup_graph_exec_, stream_));
            ++g_launches;
            led_.add(setup_graph_led_);
        } else {
            if (setup_graph_exec_) {
                cudaGraphExecDestroy(setup_graph_exec_);
                setup_graph_exec_ = nullptr;
            }
            const long long l0 = g_launches.load();
            const int saved_profile = profile_;
            profile_ = 0;  // no event records inside a capture
            cudaGraph_t graph = nullptr;
            TM_CUDA(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal));
            CaptureLedger cap(led_, capturing_);
            try {
                setup_device_part(xi);
            } catch (...) {
                cap.finish();
                cudaStreamEndCapture(stream_, &graph);
                if (graph) cudaGraphDestroy(graph);
                profile_ = saved_profile;
                throw;
            }
            setup_graph_led_ = cap.finish();
            TM_CUDA(cudaStreamEndCapture(stream_, &graph));
            profile_ = saved_profile;
            g_launches.store(l0);
            TM_CUDA(cudaGraphInstantiate(&setup_graph_exec_, graph, 0));
            cudaGraphDestroy(graph);
            setup_graph_xi_ = xi;
            TM_CUDA(cudaGraphLaunch(setup_graph_exec_, stream_));
            ++g_launches;
            led_.add(setup_graph_led_);
        }
        TM_CUDA(cudaMemcpyAsync(h_sc_ + SC_COUNT, eig_sc_, sizeof(double) * 2 * (nl - 1),
                                cudaMemcpyDeviceToHost, stream_));
        TM_CUDA(cudaStreamSynchronize(stream_));
        for (int l = 0; l + 1 < nl; ++l) {
            // after normalising by the previous norm, ||eig||^2 -> lambda^2
            const double lam = std::sqrt(h_sc_[SC_COUNT + 2 * l + 1]);
            if (!(lam > 0.0) || !(lam == lam)) throw Invalid{"multigrid: eigenvalue estimate failed"};
            levels_[l].lmax = lam;
        }
        upload_tail();
    }

    void setup_device_part(T* xi) {
        const int nl = nlevels_;
        compute_w0(xi);
        levels_[0].g.xi = xi;
        levels_[0].g.W = general_ ? w0_.p : nullptr;
        for (int l = 1; l < nl; ++l) {
            Level& F = levels_[l - 1];
            Level& C = levels_[l];
            const bool gather = nranks_ > 1 && l == dist_levels_;
            const int cell_off = C.sharded ? ranges_[l].cl0 : 0;
            int own0 = 0, own1 = C.g.ny;
            if (C.sharded || gather) {
                const RowRange rc = row_range(l, rank_, false);  // my share even if C is replicated
                own0 = rc.c0 - cell_off;
                own1 = rc.c1 - cell_off;
            }
            dim3 blk(32, 8), grd(ceil_div(C.g.nx, 32), ceil_div(C.g.ny, 8));
            if (F.g.W == nullptr)
                mg_coarsen_moments_kernel<T, false><<<grd, blk, 0, stream_>>>(F.g, C.g.nx, C.g.ny, cell_off, own0,
                                                                             own1, co_tab_, C.W.p);
            else
                mg_coarsen_moments_kernel<T, true><<<grd, blk, 0, stream_>>>(F.g, C.g.nx, C.g.ny, cell_off, own0,
                                                                            own1, co_tab_, C.W.p);
            TM_CHECK_LAUNCH();
            acct(LC_SETUP_COARSEN, (F.g.W ? 12.0 * F.g.nx * F.g.ny : (double)(F.g.nx + 1) * (F.g.ny + 1)) * sizeof(T) +
                                       12 * sz((size_t)C.g.nx * (own1 - own0)));
            if (C.sharded) exchange_w(l, C.W.p);
            if (gather) {
                const size_t plane = (size_t)C.g.nx * C.g.ny;
                for (int k = 0; k < 12; ++k) gather_rows(l, C.W.p + k * plane, (size_t)C.g.nx, true);
            }
        }
        assemble_tail();
        for (int l = 0; l + 1 < nl; ++l) {
            Level& L = levels_[l];
            launch_diag(L.g, l > 0, L.dinv.p);
            exchange_p2(l, L.dinv.p);  // the fused first smoothing step reads D^-1 on halo rows
            // lambda_max(D^-1 A) by power iteration, warm-started across solves
            const int g1 = grid1d(L.cnt);
            int its = 4;
            if (!L.eig_ready) {
                mg_seed_vector_kernel<T><<<grid1d(L.nu / 2), kVecThreads, 0, stream_>>>(L.g, L.eig.p);
                TM_CHECK_LAUNCH();
                acct(LC_SETUP_EIG, sz(L.nu));
                its = eig_first_its_;  // the top of the spectrum is clustered: the power method is slow
                L.eig_ready = true;
            }
            double* slot = eig_sc_ + 2 * l + 1;
            dot_kernel<T><<<g1, kVecThreads, 0, stream_>>>(L.cnt, L.eig.p + L.off, L.eig.p + L.off, rs_, slot);
            TM_CHECK_LAUNCH();
            acct(LC_SETUP_EIG, sz(L.cnt));
            if (L.sharded) sum_ranks(slot, 1);
            for (int it = 0; it < its; ++it) {
                exchange_p2(l, L.eig.p);
                ApplyArgs<T> a = apply_args();
                a.x = L.eig.p; a.y = L.tmp.p;
                launch_apply(L.g, l > 0, EP_PLAIN, a);
                // eig <- dinv .* tmp / ||eig_old||
                normalize_scale_kernel<T><<<g1, kVecThreads, 0, stream_>>>(
                    L.cnt, L.dinv.p + L.off, L.tmp.p + L.off, L.eig.p + L.off, slot);
                TM_CHECK_LAUNCH();
                acct(LC_SETUP_EIG, 3 * sz(L.cnt));
                dot_kernel<T><<<g1, kVecThreads, 0, stream_>>>(L.cnt, L.eig.p + L.off, L.eig.p + L.off, rs_, slot);
                TM_CHECK_LAUNCH();
                acct(LC_SETUP_EIG, sz(L.cnt));
                if (L.sharded) sum_ranks(slot, 1);
            }
        }
        {
            Level& C = levels_[nl - 1];
            // dense inverse of the coarsest operator: assembled and inverted in shared memory by one block
            coarse_Ainv_.ensure(C.nu * C.nu);
            mg_coarse_inverse_kernel<T><<<1, kCoarseInvThreads, coarse_inv_smem_, stream_>>>(C.g, coarse_Ainv_.p);
            TM_CHECK_LAUNCH();
            acct(LC_COARSE_SOLVE, 8.0 * C.nu * C.nu + 12.0 * sizeof(T) * C.g.nx * C.g.ny);
        }
    }

    // ---- mixed precision: an fp32 twin engine owns the multigrid hierarchy
    void prepare_inner(const T* xi) {
        if (!inner_) {
            tm_config c = cfg_;
            c.dtype = TM_F32;
            inner_.reset(new Engine<float>(c));
            inner_->comm_ = comm_;
            inner_->owns_comm_ = false;
            inner_->p2p_ = p2p_;
        }
        Engine<float>& in = *inner_;
        in.stream_ = stream_;
        in.cheb_degree_ = cheb_degree_; in.coarse_degree_ = coarse_degree_; in.cheb_ratio_ = cheb_ratio_;
        in.eig_safety_ = eig_safety_; in.apply_minb_ = apply_minb_; in.apply_prefetch_ = apply_prefetch_;
        in.blocks_per_sm_target_ = blocks_per_sm_target_; in.min_rows_per_strip_ = min_rows_per_strip_;
        in.use_graph_ = use_graph_; in.eig_first_its_ = eig_first_its_; in.profile_ = profile_;
        in.fuse_first_ = fuse_first_;
        in.restrict_tiled_ = restrict_tiled_;
        in.wave_aware_ = wave_aware_;
        in.fuse_cheb0_ = fuse_cheb0_;
        in.prolong_tiled_ = prolong_tiled_;
        in.cycle_first_ = cycle_first_; in.cycle_last_ = cycle_last_; in.cycle_gamma_ = cycle_gamma_;
        in.light_levels_ = light_levels_;
        in.set_penalty(spec_.p);
        in.fuse_rz_ = 0;  // r . z is taken in fp64 on the converted vectors
        if (in.tail_max_nodes_ != tail_max_nodes_ || in.tail_cluster_ != tail_cluster_) in.levels_.clear();
        in.tail_max_nodes_ = tail_max_nodes_; in.tail_cluster_ = tail_cluster_;
        in.stats_fine_applies_ = 0;
        in.stats_vcycles_ = 0;
        xi32_.ensure(n1_);
        convert_kernel<T, float><<<grid1d(n1_), kVecThreads, 0, stream_>>>(n1_, xi, xi32_.p);
        TM_CHECK_LAUNCH();
        acct(LC_CONVERT, (sizeof(T) + 4.0) * n1_);
        if (in.levels_.empty()) in.build_levels();
        in.setup_hierarchy(xi32_.p);
        in.s_r_.ensure(nu_);
        s_z_.ensure(nu_);
    }
    T* mixed_vcycle(T* r) {
        Engine<float>& in = *inner_;
        const int g1 = grid1d(p2_cnt_);
        convert_kernel<T, float><<<g1, kVecThreads, 0, stream_>>>(p2_cnt_, r + p2_off_, in.s_r_.p + p2_off_);
        TM_CHECK_LAUNCH();
        acct(LC_CONVERT, (sizeof(T) + 4.0) * p2_cnt_);
        float* z = in.vcycle(in.s_r_.p);
        convert_kernel<float, T><<<g1, kVecThreads, 0, stream_>>>(p2_cnt_, z + p2_off_, s_z_.p + p2_off_);
        TM_CHECK_LAUNCH();
        acct(LC_CONVERT, (sizeof(T) + 4.0) * p2_cnt_);
        return s_z_.p;
    }

    // Chebyshev-Jacobi smoothing of A x = b on level l; xin == nullptr means zero initial guess
    // rz_out != nullptr: the last step also delivers b . x_out there (EP_CHEBDOT)
    T* smooth(int l, const T* b, T* xin, double* rz_out = nullptr) {
        Level& L = levels_[l];
        const double hi = eig_safety_ * L.lmax, lo = hi / cheb_ratio_;
        const double theta = 0.5 * (hi + lo), delta = 0.5 * (hi - lo), sigma = theta / delta;
        double rho = 1.0 / sigma;
        T* cur;
        int k0 = 0;
        const int degree = level_degree(l);
        if (!xin && degree >= 2 && fuse_first_ && fuse_cheb0_) {
            // steps 0 and 1 from the zero guess in ONE operator pass (EP_CHEB0): x1 = D^-1 b / theta is formed
            // on the fly at the nodes the operator touches, so b's halo rows must be current
            const double rho_new = 1.0 / (2.0 * sigma - rho);
            exchange_p2(l, const_cast<T*>(b));
            ApplyArgs<T> a = apply_args();
            a.y = L.x.p; a.b = b; a.dinv = L.dinv.p; a.d = L.d.p;
            a.c0 = (T)(1.0 / theta);
            a.c1 = (T)(rho_new * rho);
            a.c2 = (T)(2.0 * rho_new / delta);
            a.store_d = degree > 2 ? 1 : 0;
            launch_apply(L.g, l > 0, EP_CHEB0, a);
            rho = rho_new;
            cur = L.x.p;
            k0 = 2;
        } else if (!xin) {
            launch_chain(cheb_first_kernel<T>, dim3(grid1d(L.cnt)), dim3(kVecThreads), L.cnt, 1.0 / theta,
                         (const T*)(L.dinv.p + L.off), b + L.off, L.d.p + L.off, L.x.p + L.off);
            TM_CHECK_LAUNCH();
            acct(LC_CHEB_FIRST, 4 * sz(L.cnt));
            cur = L.x.p;
            k0 = 1;
        } else {
            cur = xin;
        }
        for (int k = k0; k < degree; ++k) {
            double c1, c2;
            if (k == 0) {
                c1 = 0.0;
                c2 = 1.0 / theta;
            } else {
                const double rho_new = 1.0 / (2.0 * sigma - rho);
                c1 = rho_new * rho;
                c2 = 2.0 * rho_new / delta;
                rho = rho_new;
            }
            T* other = (cur == L.x.p) ? L.xalt.p : L.x.p;
            exchange_p2(l, cur);
            ApplyArgs<T> a = apply_args();
            a.x = cur; a.y = other; a.b = b; a.dinv = L.dinv.p; a.d = L.d.p;
            a.c1 = (T)c1; a.c2 = (T)c2;
            a.store_d = (k + 1 < degree) ? 1 : 0;
            const bool with_dot = rz_out != nullptr && k + 1 == degree;
            a.dot_out = rz_out;
            launch_apply(L.g, l > 0, with_dot ? EP_CHEBDOT : EP_CHEB, a);
            if (with_dot) vcycle_rz_ = true;
            cur = other;
        }
        return cur;
    }

    // z = V(r).  The ~80 small dependent launches of one V-cycle are captured once per solve
    // into a CUDA graph and replayed for the remaining PCG iterations (single rank; the sharded
    // path keeps stream launches because of its NCCL calls).  With TM_OPT_PROFILE on, the first
    // V-cycle of every solve runs un-captured so its level-0 launches can be event-timed.
    T* vcycle(T* r) {
        if (!use_graph_ || (nranks_ > 1 && !graph_sharded_) || profile_ >= 2) return vcycle_body(r);
        if (graph_exec_ && graph_r_ == r && !graph_dirty_) {
            TM_CUDA(cudaGraphLaunch(graph_exec_, stream_));
            ++g_launches;
            ++stats_vcycles_;
            stats_fine_applies_ += graph_fine_applies_;
            for (int e = 0; e < 4; ++e) fine_ep_count_[e] += graph_ep_count_[e];
            led_.add(graph_led_);
            vcycle_rz_ = graph_rz_;
            return graph_z_;
        }
        if (profile_ == 1 && !graph_sampled_) {
            graph_sampled_ = true;  // event-timed sample, capture on the next call
            return vcycle_body(r);
        }
        if (graph_exec_) {
            cudaGraphExecDestroy(graph_exec_);
            graph_exec_ = nullptr;
        }
        const int saved_profile = profile_;
        const long fa0 = stats_fine_applies_, vc0 = stats_vcycles_;
        long ep0[4];
        for (int e = 0; e < 4; ++e) ep0[e] = fine_ep_count_[e];
        const long long l0 = g_launches.load();
        profile_ = 0;
        cudaGraph_t graph = nullptr;
        TM_CUDA(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal));
        CaptureLedger cap(led_, capturing_);
        T* z = nullptr;
        try {
            z = vcycle_body(r);
        } catch (...) {
            cap.finish();
            cudaStreamEndCapture(stream_, &graph);
            if (graph) cudaGraphDestroy(graph);
            profile_ = saved_profile;
            throw;
        }
        graph_led_ = cap.finish();
        TM_CUDA(cudaStreamEndCapture(stream_, &graph));
        profile_ = saved_profile;
        graph_fine_applies_ = stats_fine_applies_ - fa0;
        stats_fine_applies_ = fa0;  // capture launched nothing
        for (int e = 0; e < 4; ++e) {
            graph_ep_count_[e] = fine_ep_count_[e] - ep0[e];
            fine_ep_count_[e] = ep0[e];
        }
        stats_vcycles_ = vc0;
        g_launches.store(l0);
        TM_CUDA(cudaGraphInstantiate(&graph_exec_, graph, 0));
        cudaGraphDestroy(graph);
        graph_r_ = r;
        graph_z_ = z;
        graph_rz_ = vcycle_rz_;
        graph_dirty_ = false;
        graph_kernels_ = 0;
        return vcycle(r);
    }

    T* vcycle_body(T* r) {
        struct PdlScope {  // programmatic dependent launches for the kernels of this V-cycle
            bool& flag;
            PdlScope(bool& f, bool on) : flag(f) { flag = on; }
            ~PdlScope() { flag = false; }
        } pdl_scope(pdl_active_, pdl_ && nranks_ == 1);
        vcycle_rz_ = false;
        int lbot = tail_first_ >= 0 ? tail_first_ : nlevels_ - 1;  // first level the recursion does not visit
        const bool truncated = depth_limit_ > 0 && depth_limit_ < lbot;  // timing studies only
        if (truncated) lbot = depth_limit_;
        T* z = cycle(0, r, nullptr, lbot, truncated);
        ++stats_vcycles_;
        return z;
    }

    // How often the coarse-grid correction of level l is repeated (options 133 / 134: levels
    // [cycle_first_, cycle_last_] are visited cycle_gamma_ times per visit of their parent: a W-cycle on
    // that window of levels, a V-cycle elsewhere).  Why: on SIMP designs (stiffness contrast 1e6 inside
    // coarse cells) the V-cycle's error grows with every level below the current one -- the scipy study
    // tools/studies/elast_smoother_study.py needs 52 PCG iterations with V against 25-27 with the
    // coarse levels cycled twice, and repeating ANY one level gains about the same, so the window sits
    // on the small levels whose visits cost microseconds.
    // Chebyshev-Jacobi steps before and after the coarse-grid correction of level l: cheb_degree_ on the finest
    // level, coarse_degree_ below -- except that levels 1..light_levels_ (option 136; automatic: levels 1-2
    // whenever a window of levels is cycled twice) take two: under the W window the PCG iteration count is set
    // by the smoothing of the small levels, not of these (bridge N=2048: 19 iterations either way, 265 -> 235 ms
    // per solve since levels 1-2 carry a third of the step's bytes; short_cantilever N=512: 24 -> 23
    // iterations, 17.8 -> 16.3 ms; profiles/r2t_cycle_study_*.jsonl).
    int level_degree(int l) const {
        if (l == 0 || coarse_degree_ <= 0) return cheb_degree_;
        int light = light_levels_ >= 0 ? light_levels_ : (repeats_any() ? 2 : 0);
        if (tail_first_ >= 0) light = std::min(light, tail_first_ - 1);
        return l <= light ? std::min(2, coarse_degree_) : coarse_degree_;
    }
    bool repeats_any() const {
        for (int l = 1; l + 1 < nlevels_; ++l)
            if (repeats(l) > 1) return true;
        return false;
    }

    int repeats(int l) const {
        if (l < 1 || l + 1 >= nlevels_) return 1;  // the coarsest level is solved exactly
        if (cycle_gamma_ > 0) return (cycle_gamma_ > 1 && l >= cycle_first_ && l <= cycle_last_) ? cycle_gamma_ : 1;
        // automatic (option 135 = 0, the default): cycle twice the levels whose short side has 8..16 cells
        // (4..32 on a bandwidth-bound mesh, where a visit of those levels is noise next to the fine level).
        // Measured on the designs real runs reach after 25 iterations (profiles/r2r_cycle_study_*.jsonl):
        // bridge N=2048 634 ms / 50 PCG iterations per solve with the V-cycle, 284 ms / 20 with levels 6-9
        // (192 x 32 ... 24 x 4 cells) cycled twice; short_cantilever N=512 25.2 ms / 46 against 18.5 ms / 24
        // with levels 5-6 (32 x 16, 16 x 8).  Wider windows gain at most two more iterations and pay for
        // them in visits of the latency-bound levels (N=512, levels 4-6: 20 iterations, 21.9 ms).
        const bool big = 2.0 * (2.0 * nx_ + 1.0) * (2.0 * nyg_ + 1.0) >= (double)((size_t)1 << 24);
        const int m = std::min(lv_nx_[l], lv_ny_[l]);
        return (m >= (big ? 4 : 8) && m <= (big ? 32 : 16)) ? 2 : 1;
    }

    // One multigrid cycle on level l for A x = b from the guess x0 (nullptr = zero): pre-smoothing,
    // coarse-grid correction (repeated repeats(l + 1) times, each visit starting from the last), post-smoothing.
    // Pre- and post-smoother are the same polynomial, so the cycle is a symmetric operator for any window.
    T* cycle(int l, const T* b, T* x0, int lbot, bool truncated) {
        const int nl = nlevels_;
        if (l == lbot) {
            if (truncated) return smooth(l, b, x0);
            if (tail_first_ >= 0) return launch_tail(x0);
            Level& C = levels_[nl - 1];
            mg_coarse_apply_inverse_kernel<T><<<1, 192, 0, stream_>>>((int)C.nu, coarse_Ainv_.p, b, C.x.p);
            TM_CHECK_LAUNCH();
            acct(LC_COARSE_SOLVE, 8.0 * C.nu * C.nu);
            return C.x.p;
        }
        Level& L = levels_[l];
        Level& C = levels_[l + 1];
        const int degree = level_degree(l);
        T* x;
        if (!x0 && degree == 1 && fuse_first_) {
            // x = (1/theta) D^-1 b and r = b - A x in ONE pass: x is formed on the fly from b and
            // D^-1 at the nodes the operator touches (saves the separate first-step kernel)
            const double hi = eig_safety_ * L.lmax, lo = hi / cheb_ratio_;
            exchange_p2(l, const_cast<T*>(b));
            ApplyArgs<T> a = apply_args();
            a.y = L.tmp.p; a.b = b; a.dinv = L.dinv.p; a.d = L.x.p;
            a.c2 = (T)(1.0 / (0.5 * (hi + lo)));
            launch_apply(L.g, l > 0, EP_RESID0, a);
            x = L.x.p;
        } else {
            x = smooth(l, b, x0);
            exchange_p2(l, x);
            ApplyArgs<T> a = apply_args();
            a.x = x; a.y = L.tmp.p; a.b = b;
            launch_apply(L.g, l > 0, EP_RESID, a);
        }
        exchange_p2(l, L.tmp.p);
        const bool gather = nranks_ > 1 && (l + 1) == dist_levels_;
        {
            dim3 blk(32, 8), grd(ceil_div(C.g.Lx, 32), ceil_div(C.g.Ly, 8));
            if (restrict_tiled_ && (long)C.g.Lx * C.g.Ly >= 16384)
                launch_chain(mg_restrict_tiled_kernel<T>, dim3(ceil_div(C.g.Lx, kRtTI), ceil_div(C.g.Ly, kRtTJ)),
                             dim3(kRtThreads), L.g, gather ? C.gpiece : C.g, (const T*)L.tmp.p, C.b.p);
            else
                launch_chain(mg_restrict_kernel<T>, grd, blk, L.g, gather ? C.gpiece : C.g, tr_tab_,
                             (const T*)L.tmp.p, C.b.p);
            TM_CHECK_LAUNCH();
            acct(LC_RESTRICT, sz(L.nu) + sz(gather ? (size_t)(C.gpiece.own_j1 - C.gpiece.own_j0) * C.g.Lx * 2 : C.cnt));
            if (gather) gather_rows(l + 1, C.b.p, (size_t)C.g.Lx * 2, false);
        }
        T* xc = cycle(l + 1, C.b.p, nullptr, lbot, truncated);
        for (int rep = 1, n = repeats(l + 1); rep < n; ++rep) xc = cycle(l + 1, C.b.p, xc, lbot, truncated);
        exchange_p2(l + 1, xc);
        {
            dim3 blk(32, 8), grd(ceil_div(L.g.Lx, 32), ceil_div(L.g.Ly, 8 * kProlongRows));
            if (prolong_tiled_ && (long)L.g.Lx * L.g.Ly >= 16384)
                launch_chain(mg_prolong_tiled_kernel<T>, dim3(ceil_div(L.g.Lx, kPtTI), ceil_div(L.g.Ly, kPtTJ)),
                             dim3(kPtThreads), L.g, C.g, (const T*)xc, x);
            else
                launch_chain(mg_prolong_add_kernel<T>, grd, blk, L.g, C.g, tr_tab_, (const T*)xc, x);
            TM_CHECK_LAUNCH();
            acct(LC_PROLONG, 2 * sz(L.cnt) + sz(C.nu));
        }
        // the outermost cycle ends with the level-0 post-smoothing: its last step also returns r . z
        const bool rz = l == 0 && (fuse_rz_ > 0 || (fuse_rz_ < 0 && p2_cnt_ >= ((size_t)1 << 24)));
        return smooth(l, b, x, rz ? sc_ + SC_RZV : nullptr);
    }

    // ------------------------------------------------------------------ state
    tm_config cfg_;
    cudaStream_t stream_ = nullptr, own_stream_ = nullptr;
    int nx_, nyg_, ny_ = 0, num_sms_ = 148;
    int rank_ = 0, nranks_ = 1, dist_levels_ = 0, nlevels_ = 1;
    NcclComm comm_ = nullptr;
    std::shared_ptr<P2PState> p2p_;  // peer windows; null = NCCL send/recv/all-reduce
    bool p2p_want_ = p2p_default();
    std::vector<int> starts_, lv_nx_, lv_ny_, lv_dr_, lv_dt_;
    std::vector<RowRange> ranges_;
    double hx_, hy_;
    size_t n1_ = 0, n2_ = 0, nu_ = 0, p2_off_ = 0, p2_cnt_ = 0, p1_off_ = 0, p1_cnt_ = 0;
    P1Geom p1_;
    LevelGeom<T> g0_;
    DiagTable diag_tab_;
    TransferTable tr_tab_;
    CoarsenTable co_tab_;
    ReduceScratch rs_{};
    double* sc_ = nullptr;
    double* eig_sc_ = nullptr;
    double* h_sc_ = nullptr;

    int precond_ = TM_PRECOND_MULTIGRID, cheb_degree_ = 1, check_every_ = 0, coarse_cells_ = 4;
    double cheb_ratio_ = 30.0, eig_safety_ = 1.1;

    DevBuf<T> f_r_, f_p_, f_Ap_, f_dinv_, f_rhs_, f_p2_;
    bool f_dinv_ready_ = false;
    DevBuf<T> s_r_, s_p_, s_Ap_, s_b_, s_dinv_, s_xi_;
    PenaltySpec spec_{3.0, 3};
    bool general_ = false;  // spec_.p != 3: level 0 runs on the stored moments w0_
    DevBuf<T> w0_;
    std::vector<Level> levels_;
    DevBuf<double> coarse_Ainv_;
    size_t coarse_inv_smem_ = 0;

    long stats_fine_applies_ = 0, stats_vcycles_ = 0;
    int profile_ = 0;
    int blocks_per_sm_target_ = 4, min_rows_per_strip_ = 1, coarse_degree_ = 3;
    bool filter_persistent_ = true, filter_cheb_ = true, filter_bounds_ready_ = false;
    std::vector<FLevel> flevels_;
    DevBuf<double> fcoarse_A_;
    bool fmg_ready_ = false;
    int eig_first_its_ = 30;
    bool mixed_ = false, owns_comm_ = true;
    std::unique_ptr<Engine<float>> inner_;
    DevBuf<float> xi32_;
    DevBuf<T> s_z_;
    int filter_mg_mode_ = 0, filter_mg_degree_ = 2;  // 0 auto, 1 multigrid, 2 never
    double filter_mg_ratio_ = 10.0;
    cudaGraphExec_t fgraph_exec_ = nullptr;
    T* fgraph_r_ = nullptr;
    T* fgraph_z_ = nullptr;
    double filter_lmin_ = 0.0, filter_lmax_ = 0.0;
    int filter_cg_iters_ = 0;
    DevBuf<double> filter_coef_;
    int apply_minb_ = 2, filter_blocks_per_sm_ = 2;
    bool use_graph_ = true, graph_dirty_ = true, graph_sampled_ = false;
    bool fuse_first_ = true;
    // option 127: r . z of the PCG from the V-cycle's last smoothing step.  -1 = automatic: on
    // bandwidth-bound meshes only (it saves two vector passes, but its grid reduction adds ~5 us to
    // the kernel, more than the separate dot costs on a latency-bound mesh: measured at N=512)
    int fuse_rz_ = -1;
    bool vcycle_rz_ = false;   // the V-cycle just run (or replayed) left r . z in sc_[SC_RZV]
    bool graph_rz_ = false;    // ... as captured in graph_exec_
    int depth_limit_ = 0, tail_dry_ = 0;
    cudaGraphExec_t setup_graph_exec_ = nullptr;
    T* setup_graph_xi_ = nullptr;
    bool filter_tb_ = true;
    bool pdl_ = true, pdl_active_ = false;  // option 126: programmatic dependent launch in V-cycles
    bool restrict_tiled_ = true;            // option 128: shared-memory tiled restriction on the large levels
    bool wave_aware_ = true;                // option 129: strip heights of the stored-moment levels fill whole waves
    bool fuse_cheb0_ = true;                // option 131: first two smoothing steps from zero in one operator pass
    bool prolong_tiled_ = true;             // option 132: shared-memory tiled prolongation on the large levels
    int light_levels_ = -1;  // option 136: levels 1..k smooth with two steps (-1: automatic, level_degree())
    int cycle_first_ = 1, cycle_last_ = 1 << 20, cycle_gamma_ = 0;  // options 133-135: W-cycle window (repeats())
    long stored_cap_ = 0;                   // resident blocks of the stored-moment operator kernel (whole GPU)
    bool warm_guard_ = true;   // option 125: drop a warm start whose residual exceeds the zero guess's
    int stats_warm_used_ = 0;  // last state solve: 1 if the caller's initial guess was kept
    int filter_tb_steps_ = 8, filter_tb_state_ = 0;  // state: 0 unplanned, 1 ready, -1 not usable
    int filter_tb_tx_ = 0, filter_tb_ty_ = 0, filter_tb_tw_ = 0, filter_tb_th_ = 0, filter_tb_rows_ = 0;
    size_t filter_tb_smem_ = 0;
    std::vector<double> filter_c12_host_;
    DevBuf<double> filter_c12_;
    double filter_c12_lmin_ = -1.0, filter_c12_lmax_ = -1.0;
    int tail_max_nodes_ = 2304, tail_cluster_ = 16, tail_cluster_used_ = 0, tail_first_ = -1;
    TailArgs<T> tail_host_;
    DevBuf<unsigned char> tail_dev_;
    DevBuf<T> tail_img_;
    size_t tail_smem_ = 0;
    bool graph_sharded_ = true;  // NCCL calls inside captured V-cycles (all ranks capture alike)
    cudaGraphExec_t graph_exec_ = nullptr;
    T* graph_r_ = nullptr;
    T* graph_z_ = nullptr;
    long graph_fine_applies_ = 0, graph_kernels_ = 0;
    long fine_ep_count_[4] = {0, 0, 0, 0}, graph_ep_count_[4] = {0, 0, 0, 0};
    bool apply_prefetch_ = true;
    int filter_blocks_ = 0;
    double* filter_part_ = nullptr;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> prof_pending_;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_free_;
    int stats_iters_ = 0;
    int expected_iters_ = 0;  // iterations of the previous converged state solve (0: unknown)
    double floor_factor_ = 0.5, stats_floor_ = 0.0, stats_rtol_used_ = 0.0;  // option 137, last state solve
    Ledger led_, graph_led_, fgraph_led_, setup_graph_led_;
    bool capturing_ = false;
    std::vector<std::pair<int, cudaEvent_t>> ev_marks_;
    std::vector<cudaEvent_t> ev_free_;
};

}  // namespace tmx

// =========================================================================================
// C ABI
// =========================================================================================
struct tm_engine_s {
    std::unique_ptr<tmx::EngineBase> impl;
};

namespace {
template <class F>
int guarded(tm_handle h, F&& f) {
    try {
        if (h) {
            cudaError_t e = cudaSetDevice(h->impl->device);
            if (e != cudaSuccess) {
                tmx::set_error(std::string("cudaSetDevice: ") + cudaGetErrorString(e));
                return TM_ERR_CUDA;
            }
        }
        f();
        return TM_OK;
    } catch (const tmx::CudaFailure& e) {
        tmx::set_error(e.what);
        return TM_ERR_CUDA;
    } catch (const tmx::Unsupported& e) {
        tmx::set_error(e.what);
        return TM_ERR_UNSUPPORTED;
    } catch (const tmx::Invalid& e) {
        tmx::set_error(e.what);
        return TM_ERR_INVALID;
    } catch (const std::exception& e) {
        tmx::set_error(e.what());
        return TM_ERR_INVALID;
    }
}
#define TM_REQUIRE_HANDLE(h)                      \
    if (!(h) || !(h)->impl) {                     \
        tmx::set_error("null engine handle");     \
        return TM_ERR_INVALID;                    \
    }
}  // namespace


// ---- single-GPU loop-back check of the peer-memory kernels (tm_p2p.cuh): R "ranks" are R streams
// of this process whose windows are plain allocations of the same device, so the hand-shake, the
// epoch bookkeeping and the memory-ordering idioms run on real hardware without a second GPU.
// The ranks run asynchronously (no host synchronisation between epochs).
namespace tmx {

// integer-valued, so that fill and check agree whatever the compiler contracts
__device__ __forceinline__ double p2p_selftest_value(int rank, int epoch, size_t i) {
    return (double)((long long)rank * 1000003LL + (long long)epoch * 7919LL + (long long)(i % 4099));
}
__global__ void p2p_selftest_fill_kernel(double* v, size_t own0_elems, size_t own_elems, int rank, int epoch) {
    TM_GRID_STRIDE(i, own_elems) v[own0_elems + i] = p2p_selftest_value(rank, epoch, i);
}
__global__ void p2p_selftest_check_kernel(const double* v, size_t row_elems, int own_rows, int up, int down, int rank,
                                          int nranks, int epoch, unsigned int* errors) {
    const size_t own_elems = (size_t)own_rows * row_elems;
    // rows from below: the neighbour's last `up` owned rows
    if (rank > 0) {
        const size_t n = (size_t)up * row_elems, src0 = own_elems - n;
        TM_GRID_STRIDE(i, n) {
            if (v[i] != p2p_selftest_value(rank - 1, epoch, src0 + i)) atomicAdd(errors, 1u);
        }
    }
    if (rank + 1 < nranks) {
        const size_t n = (size_t)down * row_elems, dst0 = (size_t)(up + own_rows) * row_elems;
        TM_GRID_STRIDE(i, n) {
            if (v[dst0 + i] != p2p_selftest_value(rank + 1, epoch, i)) atomicAdd(errors, 1u);
        }
    }
}
__global__ void p2p_selftest_red_fill_kernel(double* p, int rank, int epoch) {
    p[0] = (rank + 1) * (double)epoch;
    p[1] = 0.5 * rank - epoch;
}
__global__ void p2p_selftest_red_check_kernel(const double* p, int nranks, int epoch, unsigned int* errors) {
    double a = 0.0, b = 0.0;
    for (int q = 0; q < nranks; ++q) {
        a += (q + 1) * (double)epoch;
        b += 0.5 * q - epoch;
    }
    if (p[0] != a || p[1] != b) atomicAdd(errors, 1u);
}

// report: [0] data mismatches, [1] timed-out polls (header.error != 0 on any rank), [2] final halo
// epoch of rank 0, [3] final reduction epoch of rank 0
static void p2p_selftest(int R, int epochs, int row_elems, int reduce_every, double* report) {
    if (R < 2 || R > 8 || epochs < 1 || row_elems < 1) throw Invalid{"tm_p2p_selftest: bad arguments"};
    const int up = 4, down = 3, own_rows = 10;
    const size_t rb = (size_t)row_elems * sizeof(double);
    const size_t slot = (((size_t)up * rb) + 255) & ~(size_t)255;
    const size_t wbytes = P2P_HEADER_BYTES + 4 * slot;
    const size_t vcount = (size_t)(up + own_rows + down) * row_elems;
    std::vector<char*> win(R, nullptr);
    std::vector<double*> v(R, nullptr), red(R, nullptr);
    std::vector<cudaStream_t> st(R, nullptr);
    unsigned int* d_err = nullptr;
    auto cleanup = [&] {
        for (int r = 0; r < R; ++r) {
            if (st[r]) cudaStreamDestroy(st[r]);
            cudaFree(win[r]);
            cudaFree(v[r]);
            cudaFree(red[r]);
        }
        cudaFree(d_err);
    };
    try {
        TM_CUDA(cudaMalloc(&d_err, sizeof(unsigned int)));
        TM_CUDA(cudaMemset(d_err, 0, sizeof(unsigned int)));
        for (int r = 0; r < R; ++r) {
            TM_CUDA(cudaMalloc(&win[r], wbytes));
            TM_CUDA(cudaMemset(win[r], 0, wbytes));
            TM_CUDA(cudaMalloc(&v[r], vcount * sizeof(double)));
            TM_CUDA(cudaMemset(v[r], 0, vcount * sizeof(double)));
            TM_CUDA(cudaMalloc(&red[r], 2 * sizeof(double)));
            TM_CUDA(cudaStreamCreateWithFlags(&st[r], cudaStreamNonBlocking));
        }
        // load every kernel now: lazily loading one while another stream's kernel polls for a kernel
        // this thread has yet to launch would stall the launch (single process, single context)
        cudaFuncAttributes fa;
        TM_CUDA(cudaFuncGetAttributes(&fa, p2p_selftest_fill_kernel));
        TM_CUDA(cudaFuncGetAttributes(&fa, p2p_selftest_check_kernel));
        TM_CUDA(cudaFuncGetAttributes(&fa, p2p_selftest_red_fill_kernel));
        TM_CUDA(cudaFuncGetAttributes(&fa, p2p_selftest_red_check_kernel));
        TM_CUDA(cudaFuncGetAttributes(&fa, p2p_halo_kernel<uint4>));
        TM_CUDA(cudaFuncGetAttributes(&fa, p2p_halo_kernel<uint2>));
        TM_CUDA(cudaFuncGetAttributes(&fa, p2p_allreduce_kernel));
        TM_CUDA(cudaDeviceSynchronize());
        const size_t own_elems = (size_t)own_rows * row_elems;
        for (int e = 1; e <= epochs; ++e) {
            // epoch-major launch order: every kernel a poll can wait for is already enqueued
            for (int r = 0; r < R; ++r) {
                p2p_selftest_fill_kernel<<<4, 256, 0, st[r]>>>(v[r], (size_t)up * row_elems, own_elems, r, e);
                TM_CHECK_LAUNCH();
                P2PHaloArgs a;
                a.self = win[r];
                a.below = r > 0 ? win[r - 1] : nullptr;
                a.above = r + 1 < R ? win[r + 1] : nullptr;
                a.slot_bytes = slot;
                a.v = reinterpret_cast<char*>(v[r]);
                a.up_src = (size_t)(up + own_rows - up) * rb;
                a.up_bytes = (size_t)up * rb;
                a.down_src = (size_t)up * rb;
                a.down_bytes = (size_t)down * rb;
                a.from_below_dst = 0;
                a.from_above_dst = (size_t)(up + own_rows) * rb;
                const int blocks = 1 + (e % 3);  // exercises the last-block bookkeeping
                if (rb % 16 == 0) p2p_halo_kernel<uint4><<<blocks, 256, 0, st[r]>>>(a);
                else p2p_halo_kernel<uint2><<<blocks, 256, 0, st[r]>>>(a);
                TM_CHECK_LAUNCH();
                p2p_selftest_check_kernel<<<4, 256, 0, st[r]>>>(v[r], (size_t)row_elems, own_rows, up, down, r, R, e,
                                                               d_err);
                TM_CHECK_LAUNCH();
                if (reduce_every > 0 && e % reduce_every == 0) {
                    p2p_selftest_red_fill_kernel<<<1, 1, 0, st[r]>>>(red[r], r, e);
                    TM_CHECK_LAUNCH();
                    P2PReduceArgs ra;
                    for (int q = 0; q < P2P_MAX_RANKS; ++q) ra.win[q] = q < R ? win[q] : nullptr;
                    ra.rank = r;
                    ra.nranks = R;
                    ra.p = red[r];
                    ra.n = 2;
                    p2p_allreduce_kernel<<<1, 64, 0, st[r]>>>(ra);
                    TM_CHECK_LAUNCH();
                    p2p_selftest_red_check_kernel<<<1, 1, 0, st[r]>>>(red[r], R, e, d_err);
                    TM_CHECK_LAUNCH();
                }
            }
        }
        for (int r = 0; r < R; ++r) TM_CUDA(cudaStreamSynchronize(st[r]));
        unsigned int err = 0, timeouts = 0;
        TM_CUDA(cudaMemcpy(&err, d_err, sizeof(err), cudaMemcpyDeviceToHost));
        P2PHeader h0;
        for (int r = 0; r < R; ++r) {
            P2PHeader h;
            TM_CUDA(cudaMemcpy(&h, win[r], sizeof(h), cudaMemcpyDeviceToHost));
            if (h.error) ++timeouts;
            if (r == 0) h0 = h;
        }
        report[0] = err;
        report[1] = timeouts;
        report[2] = (double)h0.halo_epoch;
        report[3] = (double)h0.red_epoch;
    } catch (...) {
        cleanup();
        throw;
    }
    cleanup();
}

}  // namespace tmx

// fluid problem: its own small driver object, see tm_fluid_cuda.cuh
struct tm_fluid_s {
    std::unique_ptr<tmx::FluidSolver> impl;
};

template <typename F>
static int fluid_guarded(tm_fluid_s* h, F&& f) {
    if (!h || !h->impl) {
        tmx::set_error("null fluid handle");
        return TM_ERR_INVALID;
    }
    return guarded(nullptr, [&] {
        TM_CUDA(cudaSetDevice(h->impl->device()));
        f();
    });
}

extern "C" {

const char* tm_last_error(void) { return tmx::last_error(); }
const char* tm_version(void) { return "topomax_b200 0.2 (sm_100a)"; }

int tm_create(const tm_config* cfg, tm_handle* out) {
    if (!cfg || !out) {
        tmx::set_error("tm_create: null argument");
        return TM_ERR_INVALID;
    }
    *out = nullptr;
    return guarded(nullptr, [&] {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw tmx::CudaFailure{std::string("no CUDA device available: ") + cudaGetErrorString(e)};
        if (cfg->device < 0 || cfg->device >= ndev) throw tmx::Invalid{"bad device ordinal"};
        auto* h = new tm_engine_s;
        try {
            if (cfg->dtype == TM_F64) h->impl.reset(new tmx::Engine<double>(*cfg));
            else if (cfg->dtype == TM_F32) h->impl.reset(new tmx::Engine<float>(*cfg));
            else throw tmx::Invalid{"dtype must be TM_F64 or TM_F32"};
        } catch (...) {
            delete h;
            throw;
        }
        *out = h;
    });
}

int tm_destroy(tm_handle h) {
    if (!h) return TM_OK;
    delete h;
    return TM_OK;
}

int tm_set_stream(tm_handle h, void* s) {
    TM_REQUIRE_HANDLE(h);
    return guarded(h, [&] { h->impl->set_stream((cudaStream_t)s); });
}
int tm_set_option(tm_handle h, int option, double value) {
    TM_REQUIRE_HANDLE(h);
    return guarded(h, [&] { h->impl->set_option(option, value); });
}
int tm_comm_unique_id(char* id128) {
    if (!id128) {
        tmx::set_error("tm_comm_unique_id: null argument");
        return TM_ERR_INVALID;
    }
    return guarded(nullptr, [&] {
        if (!tmx::nccl().load()) throw tmx::Invalid{tmx::nccl().error};
        tmx::NcclUniqueId id;
        const int rc = tmx::nccl().GetUniqueId(&id);
        if (rc != 0) throw tmx::Invalid{std::string("ncclGetUniqueId: ") + tmx::nccl().GetErrorString(rc)};
        std::memcpy(id128, id.internal, 128);
    });
}
int tm_comm_init(tm_handle h, const char* id128) {
    TM_REQUIRE_HANDLE(h);
    if (!id128) {
        tmx::set_error("tm_comm_init: null argument");
        return TM_ERR_INVALID;
    }
    return guarded(h, [&] { h->impl->comm_init(id128); });
}
int tm_local_layout(tm_handle h, int* out, int n) {
    TM_REQUIRE_HANDLE(h);
    return guarded(h, [&] { h->impl->layout(out, n); });
}
int tm_load_vector(tm_handle h, const tm_loads* loads, void* b) {
    TM_REQUIRE_HANDLE(h);
    if (!loads || !b) {
        tmx::set_error("tm_load_vector: null argument");
        return TM_ERR_INVALID;
    }
    return guarded(h, [&] { h->impl->load_vector(*loads, b); });
}
int tm_filter_apply(tm_handle h, int rhs_kind, void* in, void* out, double rtol, int maxit, int* iters,
                    double* relres) {
    TM_REQUIRE_HANDLE(h);
    tmx::SolveStats st;
    int rc = guarded(h, [&] { st = h->impl->filter_apply(rhs_kind, in, out, rtol, maxit); });
    if (iters) *iters = st.iters;
    if (relres) *relres = st.relres;
    if (rc == TM_OK && !st.converged) {
        tmx::set_error("filter PCG did not converge: relres " + std::to_string(st.relres));
        return TM_ERR_NOT_CONVERGED;
    }
    return rc;
}
int tm_elast_matvec(tm_handle h, void* xi, double penalty, void* x, void* y) {
    TM_REQUIRE_HANDLE(h);
    return guarded(h, [&] { h->impl->elast_matvec(xi, penalty, x, y); });
}
int tm_elast_diag(tm_handle h, void* xi, double penalty, void* dinv) {
    TM_REQUIRE_HANDLE(h);
    return guarded(h, [&] { h->impl->elast_diag(xi, penalty, dinv); });
}
int tm_state_solve(tm_handle h, void* xi, double penalty, const void* b, void* u, double rtol, int maxit,
                   int flags, int* iters, double* relres) {
    TM_REQUIRE_HANDLE(h);
    tmx::SolveStats st;
    int rc = guarded(h, [&] { st = h->impl->state_solve(xi, penalty, b, u, rtol, maxit, flags); });
    if (iters) *iters = st.iters;
    if (relres) *relres = st.relres;
    if (rc == TM_OK && !st.converged) {
        tmx::set_error("state PCG did not converge: relres " + std::to_string(st.relres));
        return TM_ERR_NOT_CONVERGED;
    }
    return rc;
}
int tm_dot_p2(tm_handle h, const void* u, const void* b, double* out) {
    TM_REQUIRE_HANDLE(h);
    return guarded(h, [&] { *out = h->impl->dot_p2(u, b); });
}
int tm_sens_rhs(tm_handle h, const void* xi, double penalty, const void* u, void* out) {
    TM_REQUIRE_HANDLE(h);
    return guarded(h, [&] { h->impl->sens_rhs(xi, penalty, u, out); });
}
int tm_md_halfstep(tm_handle h, const void* psi, const void* grad, double alpha, void* half) {
    TM_REQUIRE_HANDLE(h);
    return guarded(h, [&] { h->impl->md_halfstep(psi, grad, alpha, half); });
}
int tm_md_volume(tm_handle h, const void* half, double c, double* vol, double* dvol) {
    TM_REQUIRE_HANDLE(h);
    return guarded(h, [&] { h->impl->md_volume(half, c, vol, dvol); });
}
int tm_md_apply(tm_handle h, const void* half, double c, const void* psi_prev, void* psi, void* rho,
                double* delta_sq, double* vol) {
    TM_REQUIRE_HANDLE(h);
    return guarded(h, [&] { h->impl->md_apply(half, c, psi_prev, psi, rho, delta_sq, vol); });
}
int tm_md_project(tm_handle h, const void* half, double volume, double tol, int maxit, double* c, int* iters,
                  int* status) {
    TM_REQUIRE_HANDLE(h);
    if (!half || !c || !iters || !status || maxit < 1) {
        tmx::set_error("tm_md_project: bad argument");
        return TM_ERR_INVALID;
    }
    return guarded(h, [&] { *status = h->impl->md_project(half, volume, tol, maxit, c, iters); });
}
int tm_sample_field(tm_handle h, int degree, const void* field, int nsx, int nsy, double x0, double dx, double y0,
                    double dy, void* out) {
    TM_REQUIRE_HANDLE(h);
    return guarded(h, [&] { h->impl->sample_field(degree, field, nsx, nsy, x0, dx, y0, dy, out); });
}

int tm_integrate(tm_handle h, const void* values, double* out) {
    TM_REQUIRE_HANDLE(h);
    return guarded(h, [&] { *out = h->impl->integrate(values); });
}
int tm_last_solve_stats(tm_handle h, double* out, int n) {
    TM_REQUIRE_HANDLE(h);
    return guarded(h, [&] { h->impl->last_stats(out, n); });
}
int tm_profile_read(tm_handle h, double* out, int n) {
    TM_REQUIRE_HANDLE(h);
    return guarded(h, [&] { h->impl->profile_read(out, n); });
}
int tm_ledger_read(tm_handle h, double* out, int n, int reset) {
    TM_REQUIRE_HANDLE(h);
    if (!out || n < 0) {
        tmx::set_error("tm_ledger_read: bad argument");
        return TM_ERR_INVALID;
    }
    return guarded(h, [&] { h->impl->ledger_read(out, n, reset); });
}
long long tm_launch_count(void) { return tmx::g_launches.load(); }
int tm_mg_debug(tm_handle h, void* xi, int op, int level, const void* in, void* out) {
    TM_REQUIRE_HANDLE(h);
    return guarded(h, [&] { h->impl->mg_debug(xi, op, level, in, out); });
}
int tm_mg_level_info(tm_handle h, int level, int* info6, int* nlevels) {
    TM_REQUIRE_HANDLE(h);
    return guarded(h, [&] { *nlevels = h->impl->mg_level_info(level, info6); });
}

// ---- fluid problem: see tm_fluid_cuda.cuh
int tm_fluid_create(int nx, int ny, double width, double height, double viscosity, double r_min, double r_max,
                    int device, tm_fluid_handle* out) {
    if (!out) {
        tmx::set_error("tm_fluid_create: null argument");
        return TM_ERR_INVALID;
    }
    *out = nullptr;
    return guarded(nullptr, [&] {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw tmx::CudaFailure{std::string("no CUDA device available: ") + cudaGetErrorString(e)};
        if (device < 0 || device >= ndev) throw tmx::Invalid{"bad device ordinal"};
        auto* h = new tm_fluid_s;
        try {
            h->impl.reset(new tmx::FluidSolver(nx, ny, width, height, viscosity, r_min, r_max, device));
        } catch (...) {
            delete h;
            throw;
        }
        *out = h;
    });
}
int tm_fluid_destroy(tm_fluid_handle h) {
    delete h;
    return TM_OK;
}
int tm_fluid_set_stream(tm_fluid_handle h, void* stream) {
    return fluid_guarded(h, [&] { h->impl->set_stream(static_cast<cudaStream_t>(stream)); });
}
int tm_fluid_set_option(tm_fluid_handle h, int option, double value) {
    return fluid_guarded(h, [&] { h->impl->set_option(option, value); });
}
int tm_fluid_set_density(tm_fluid_handle h, const double* rho, double q) {
    return fluid_guarded(h, [&] {
        if (!rho) throw tmx::Invalid{"tm_fluid_set_density: null argument"};
        h->impl->set_density(rho, q);
    });
}
int tm_fluid_state_solve(tm_fluid_handle h, const double* boundary_velocity, double rtol, int maxit, double* up,
                         int* iters, double* relres) {
    int rc_extra = TM_OK;
    const int rc = fluid_guarded(h, [&] {
        if (!boundary_velocity || !up) throw tmx::Invalid{"tm_fluid_state_solve: null argument"};
        const tmx::MinresResult r = h->impl->solve(boundary_velocity, rtol, maxit, up);
        if (iters) *iters = r.iterations;
        if (relres) *relres = r.relres;
        if (!r.converged) {
            tmx::set_error("tm_fluid_state_solve: MINRES did not reach the tolerance (relative residual " +
                           std::to_string(r.relres) + " after " + std::to_string(r.iterations) + " iterations)");
            rc_extra = TM_ERR_NOT_CONVERGED;
        }
    });
    return rc != TM_OK ? rc : rc_extra;
}
int tm_fluid_objective(tm_fluid_handle h, const double* u, double* out) {
    return fluid_guarded(h, [&] {
        if (!u || !out) throw tmx::Invalid{"tm_fluid_objective: null argument"};
        *out = h->impl->objective(u);
    });
}
int tm_fluid_sens_rhs(tm_fluid_handle h, const double* rho, const double* u, double* out) {
    return fluid_guarded(h, [&] {
        if (!rho || !u || !out) throw tmx::Invalid{"tm_fluid_sens_rhs: null argument"};
        h->impl->sens_rhs(rho, u, out);
        h->impl->sync();
    });
}
int tm_fluid_apply(tm_fluid_handle h, const double* x, double* y, int mode) {
    return fluid_guarded(h, [&] {
        if (!x || !y || (mode != 0 && mode != 1)) throw tmx::Invalid{"tm_fluid_apply: bad argument"};
        h->impl->apply_mode(x, y, mode);
        h->impl->sync();
    });
}

int tm_p2p_selftest(int nranks, int epochs, int row_elems, int reduce_every, double* report4) {
    return guarded(nullptr, [&] {
        if (!report4) throw tmx::Invalid{"tm_p2p_selftest: null argument"};
        tmx::p2p_selftest(nranks, epochs, row_elems, reduce_every, report4);
    });
}

// Stateless: the evaluator belongs to no mesh hierarchy.  Scratch for the deterministic
// reduction is allocated per call (the arrays of this path are small).
int tm_dem_strain_energy(int nx, int ny, double width, double height, double lame_lambda, double lame_mu,
                         double simp_min, double penalty, const float* u, const float* density,
                         float* cell_energy, float* grad_density, float* grad_u, double* objective,
                         void* stream) {
    return guarded(nullptr, [&] {
        if (nx < 1 || ny < 1 || !(width > 0) || !(height > 0)) throw tmx::Invalid{"tm_dem_strain_energy: bad mesh"};
        if (!u || !density || !objective) throw tmx::Invalid{"tm_dem_strain_energy: null argument"};
        const tmx::DemGeom g = tmx::dem_make_geom(nx, ny, width, height, lame_lambda, lame_mu, simp_min, penalty);
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const size_t cells = (size_t)nx * ny, nodes = (size_t)(nx + 1) * (ny + 1);
        int sms = 148;
        cudaPointerAttributes attr;
        TM_CUDA(cudaPointerGetAttributes(&attr, u));
        if (attr.type != cudaMemoryTypeDevice) throw tmx::Invalid{"tm_dem_strain_energy: u is not device memory"};
        const int dev = attr.device;
        TM_CUDA(cudaSetDevice(dev));
        TM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int blocks = (int)std::min<size_t>((cells + 255) / 256, (size_t)sms * 8);
        tmx::DevBuf<double> scratch;  // [0] objective, [1..] block partials; counter behind them
        scratch.ensure((size_t)blocks + 4);
        tmx::ReduceScratch rs;
        rs.partials = scratch.p + 2;
        rs.capacity = blocks;
        rs.counter = reinterpret_cast<unsigned int*>(scratch.p + 1);
        TM_CUDA(cudaMemsetAsync(scratch.p, 0, 2 * sizeof(double), st));  // ordered on the caller's stream
        tmx::dem_cell_kernel<<<blocks, 256, 0, st>>>(g, u, density, cell_energy, grad_density, rs, scratch.p);
        TM_CHECK_LAUNCH();
        if (grad_u) {
            const int nb = (int)std::min<size_t>((nodes + 255) / 256, (size_t)sms * 8);
            tmx::dem_grad_u_kernel<<<nb, 256, 0, st>>>(g, u, density, grad_u);
            TM_CHECK_LAUNCH();
        }
        TM_CUDA(cudaMemcpyAsync(objective, scratch.p, sizeof(double), cudaMemcpyDeviceToHost, st));
        TM_CUDA(cudaStreamSynchronize(st));
    });
}

}  // extern "C"
