// bdd_b200/csrc/bdd_b200.cu -- implementation of the C ABI in include/bdd_b200.h.
//
// Host-side state machine mirrors bdd_cuda_base<REAL> / bdd_cuda_parallel_mma<REAL>
// (forward_state_valid_ / backward_state_valid_, include/bdd_solver/bdd_cuda_base.h:211-212);
// all device work is the kernels in kernels.cuh on one CUDA stream.  No thrust, no
// per-hop launches, no host synchronisation except where a result is returned to the host.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <cstring>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

#include "kernels.cuh"
#include "resident.cuh"
#include "shard.hpp"
#include "last_error.hpp"

using namespace bddb200;

namespace {

thread_local std::string g_last_error;

struct cuda_error : std::runtime_error {
    cuda_error(const std::string& m) : std::runtime_error(m) {}
};
struct api_error : std::runtime_error {
    int code;
    api_error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define CUDA_CHECK(expr) do { cudaError_t e_ = (expr); if(e_ != cudaSuccess) \
    throw cuda_error(std::string(#expr) + ": " + cudaGetErrorString(e_)); } while(0)

template<typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if(p) cudaFree(p); }
    void alloc(size_t count) { if(p) { cudaFree(p); p = nullptr; } n = count; CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T))); }
    template<typename A>
    void upload(const std::vector<T, A>& h, cudaStream_t st) { alloc(h.size()); if(!h.empty()) CUDA_CHECK(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st)); }
    void zero(cudaStream_t st) { if(n) CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(T), st)); }
    void clone_from(const DevBuf& o, cudaStream_t st) { alloc(o.n); if(o.n) CUDA_CHECK(cudaMemcpyAsync(p, o.p, o.n * sizeof(T), cudaMemcpyDeviceToDevice, st)); }
};

inline unsigned blocks_for(size_t n, unsigned t = 256) { return (unsigned)((n + t - 1) / t); }

} // namespace

// ---------------------------------------------------------------------------------------
struct bddb200_solver {
    int precision = 0;
    int device = 0;
    virtual ~bddb200_solver() {}
    virtual size_t nr_variables() const = 0;
    virtual size_t nr_bdds() const = 0;
    virtual size_t nr_layers() const = 0;
    virtual size_t nr_bdd_nodes() const = 0;
    virtual size_t nr_hops() const = 0;
    virtual void nr_bdds_per_var(int32_t* out) const = 0;
    virtual void layer_primal_indices(int32_t* out) const = 0;
    virtual void layer_bdd_indices(int32_t* out) const = 0;
    virtual void iteration(double omega) = 0;
    virtual void iterations(double omega, size_t n) = 0;
    virtual void forward_pass(double omega) = 0;
    virtual void backward_pass(double omega) = 0;
    virtual void forward_mm(double omega, void* delta_dev) = 0;
    virtual void backward_mm(double omega, void* delta_dev) = 0;
    virtual void normalize_delta(void* delta_dev) const = 0;
    virtual void get_delta(void* out, int out_is_host) = 0;
    virtual double lower_bound() = 0;
    virtual void lower_bound_per_bdd(void* out_dev) = 0;
    virtual void forward_run() = 0;
    virtual void backward_run() = 0;
    virtual void flush_forward() = 0;
    virtual void flush_backward() = 0;
    virtual void update_costs_host(const double* lo, size_t n_lo, const double* hi, size_t n_hi) = 0;
    virtual void update_costs_host_real(const void* lo, size_t n_lo, const void* hi, size_t n_hi) = 0;
    virtual void update_costs_dev(const void* lo, size_t n_lo, const void* hi, size_t n_hi) = 0;
    virtual double step_host(const void* lo, size_t n_lo, const void* hi, size_t n_hi, int src_is_real, double omega) = 0;
    virtual void set_cost(double c, size_t var) = 0;
    virtual void distribute_delta() = 0;
    virtual void get_solver_costs(void* lo, void* hi, void* mmd) const = 0;
    virtual void set_solver_costs(const void* lo, const void* hi, const void* mmd) = 0;
    virtual void primal_objective_host(double* out) = 0;
    virtual void min_marginals(int sorted, int32_t* primal_dev, void* lo_dev, void* hi_dev) = 0;
    virtual void min_marginals_host(int sorted, int32_t* primal_host, double* lo_host, double* hi_host) = 0;
    virtual void bdds_solution(char* sol_dev) = 0;
    virtual void net_solver_costs(void* out_dev) const = 0;
    virtual void make_dual_feasible(void* inout_dev) const = 0;
    virtual void gradient_step(const void* dir_dev, double step) = 0;
    virtual void synchronize() = 0;
    virtual void* stream_handle() = 0;
    virtual size_t kernel_launches() const = 0;
    virtual void* delta_sum_buffer() = 0;
    virtual int rounding_perturb(double delta, int round_index, unsigned long long counts_out[4], char* types_dev, char* sol_host) = 0;
    virtual int delta_sum_index() const = 0;
    virtual int push_exchange_supported() const = 0;
    virtual void set_push_masks(const uint16_t* masks_host, size_t n) = 0;
    virtual void set_delta_buffers(void* b0, void* b1, void* b2) = 0;
    virtual void set_delta_input(void* in, size_t n_shared_vars) = 0;
    virtual void set_exchange(int world, int rank, const void* const* peers, uint32_t* const* flags, void* out, void* const* outs,
                              const void* mc_in, void* mc_out, size_t n_exchange, int mode) = 0;
    virtual size_t trace_pass(int forward, double omega, unsigned long long* out_host, size_t max_bundles) = 0;
    virtual bddb200_solver* clone() const = 0;
    virtual void save(std::vector<unsigned char>& out) const = 0;
};

#include "lbfgs.cuh"

namespace {

template<typename REAL>
class SolverImpl final : public bddb200_solver {
public:
    SolverImpl(const bddb200_instruction* instrs, size_t n_instr, const size_t* delims, size_t n_bdds,
               const double* costs_hi, size_t n_costs, const bddb200_options& opt)
    {
        precision = sizeof(REAL) == 8 ? BDDB200_DOUBLE : BDDB200_FLOAT;
        device = opt.device;
        deterministic_ = opt.deterministic != 0;
        int count = 0;
        if(cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
            throw api_error(BDDB200_ERR_NO_DEVICE, "no CUDA device available: libbdd_b200 has no CPU fallback");
        if(device < 0 || device >= count) throw api_error(BDDB200_ERR_INVALID_ARGUMENT, "invalid device ordinal");
        CUDA_CHECK(cudaSetDevice(device));
        if(opt.stream) { stream_ = (cudaStream_t)opt.stream; own_stream_ = false; }
        else { CUDA_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking)); own_stream_ = true; }

        size_t stage_budget = opt.stage_bytes ? (size_t)opt.stage_bytes : DEFAULT_STAGE_BUDGET;
        if(const char* e = std::getenv("BDDB200_STAGE_BYTES")) stage_budget = (size_t)std::atol(e);
        n_stages_ = opt.n_stages ? (unsigned)opt.n_stages : 3u;
        if(const char* e = std::getenv("BDDB200_STAGES")) n_stages_ = (unsigned)std::atoi(e);
        if(n_stages_ < 2 || n_stages_ > 8) throw api_error(BDDB200_ERR_INVALID_ARGUMENT, "n_stages must be in [2, 8]");
        if(stage_budget < 1024 || stage_budget > 96 * 1024) throw api_error(BDDB200_ERR_INVALID_ARGUMENT, "stage_bytes must be in [1 KiB, 96 KiB]");
        forced_wpc_ = opt.warps_per_cta > 0 ? (unsigned)opt.warps_per_cta : 0u;
        if(const char* e = std::getenv("BDDB200_WARPS_PER_CTA")) forced_wpc_ = (unsigned)std::atoi(e);
        if(forced_wpc_ > 16) throw api_error(BDDB200_ERR_INVALID_ARGUMENT, "warps_per_cta must be <= 16");

        int lanes_per_bdd = opt.lanes_per_bdd;
        if(const char* e = std::getenv("BDDB200_LANES_PER_BDD")) lanes_per_bdd = std::atoi(e);
        bool lane_class = true;       // BDDB200_NO_LANE=1: run narrow one-lane-per-BDD bundles through the generic kernel (A/B checks)
        if(const char* e = std::getenv("BDDB200_NO_LANE")) lane_class = std::atoi(e) == 0;
        CUDA_CHECK(cudaDeviceGetAttribute(&n_sms_, cudaDevAttrMultiProcessorCount, device));
        size_t balance_sms = 0;      // BDDB200_BALANCE=1: bundle count a multiple of the SM count, BDDs dealt evenly (measured neutral on B200: the exchange is bound chip-wide, not per SM)
        if(const char* e = std::getenv("BDDB200_BALANCE")) if(std::atoi(e) != 0) balance_sms = (size_t)n_sms_;
        HostLayout L = build_layout(instrs, n_instr, delims, n_bdds, lanes_per_bdd, opt.nr_variables, sizeof(REAL), stage_budget, lane_class, balance_sms, opt.n_shared_vars);
        n_lane_ = L.n_lane_bundles; lane_max_J_ = L.lane_max_J; lane_max_hops_ = L.lane_max_hops;
        n_lane_shared_ = L.n_lane_shared_bundles; layout_shared_vars_ = opt.n_shared_vars;
        {   // runs of lane-class bundles with equal (J, n_hops): arithmetic progressions in every array (layout.hpp emits them back to back)
            bool ok = std::getenv("BDDB200_NO_CLASS_DESC") == nullptr;
            for(size_t g = 0; g < L.desc_lane.size() && ok; ++g)
            {
                const LaneDesc& d = L.desc_lane[g];
                if(!lane_cls_begin_.empty())
                {
                    const LaneDesc& f = lane_cls_first_.back();
                    const uint32_t q = (uint32_t)g - lane_cls_begin_.back();
                    if(d.J == f.J && d.n_hops == f.n_hops && d.slot_off == f.slot_off + q * f.n_hops * f.J * 32u && d.lay_off == f.lay_off + q * f.n_hops * 32u
                       && d.topo_off == f.topo_off + q * f.n_hops * 32u && d.bdd_base == f.bdd_base + q * 32u) continue;
                }
                if(lane_cls_begin_.size() == (size_t)LANE_MAX_CLASSES) { ok = false; break; }
                lane_cls_first_.push_back(d); lane_cls_begin_.push_back((uint32_t)g);
            }
            if(!ok) { lane_cls_first_.clear(); lane_cls_begin_.clear(); }
        }
        n_vars_ = L.n_vars; n_bdds_ = L.n_bdds; n_instr_ = delims[n_bdds] - delims[0];
        n_ext_ = L.n_layers_ext; n_slots_ = L.n_slots; n_lay_ = L.n_lay; max_hops_ = L.max_hops;
        n_bundles_ = L.bundles.size(); n_small_ = L.n_small_bundles;
        tile_small_ = std::max<uint32_t>(L.max_tile_small, 32u); tile_large_ = L.max_tile_large;
        stage_small_ = (uint32_t)((std::max<size_t>(L.stage_small, 128) + 127) & ~(size_t)127);
        stage_large_ = (uint32_t)((L.stage_large + 127) & ~(size_t)127);

        h_nr_bdds_per_var_ = L.nr_bdds_per_var;
        if(opt.nr_bdds_per_var_host != nullptr)
        {
            if(opt.nr_variables < L.n_vars) throw api_error(BDDB200_ERR_INVALID_ARGUMENT, "nr_bdds_per_var_host given but nr_variables too small");
            for(size_t v = 0; v < n_vars_; ++v)
            {
                if(opt.nr_bdds_per_var_host[v] < h_nr_bdds_per_var_[v])
                    throw api_error(BDDB200_ERR_INVALID_ARGUMENT, "global nr_bdds_per_var smaller than the local count");
                h_nr_bdds_per_var_[v] = opt.nr_bdds_per_var_host[v];
            }
        }
        h_ext_var_.assign(L.ext_var.begin(), L.ext_var.end()); h_ext_bdd_.assign(L.ext_bdd.begin(), L.ext_bdd.end());

        // the lane-class kernels normalise through a reciprocal table; a variable shared by more BDDs than it holds
        // switches the solver to exact division + fixed-order sums (the deterministic kernels)
        if(n_lane_ > 0 && !deterministic_ && !h_nr_bdds_per_var_.empty() && *std::max_element(h_nr_bdds_per_var_.begin(), h_nr_bdds_per_var_.end()) >= INV_TAB)
        {
            deterministic_ = true;
            if(std::getenv("BDDB200_QUIET") == nullptr)
                std::fprintf(stderr, "[bdd_b200] a variable occurs in >= %d BDDs: using the deterministic kernels (exact division, fixed-order sums)\n", INV_TAB);
        }
        CUDA_CHECK(cudaDeviceGetAttribute(&max_optin_, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        CUDA_CHECK(cudaDeviceGetAttribute(&n_sms_, cudaDevAttrMultiProcessorCount, device));
        plan_launch();
        // a collection of few, long BDDs leaves most of the GPU idle: a pass is then one dependency chain of max_hops steps per warp
        // (assignment_5m, 2 x 1118 hops: 0.17 of the HBM roofline; after split_qbdd with chunk length 64: 0.68).  Tell the caller once.
        if(n_bundles_ < (size_t)n_sms_ * 8 && max_hops_ >= 256 && std::getenv("BDDB200_QUIET") == nullptr)
            std::fprintf(stderr, "[bdd_b200] %zu bundles of up to %zu hops occupy fewer than 8 warps on each of the %d SMs: every pass is a chain of %zu dependent steps; "
                                 "splitting the long BDDs (bdd_collection::split_qbdd; driver key \"split bdds\") shortens it\n", n_bundles_, max_hops_, n_sms_, max_hops_);
        L_var_lay_begin_ = L.var_lay_begin;
        plan_resident();
        L_var_lay_begin_.clear(); L_var_lay_begin_.shrink_to_fit();

        std::vector<uint32_t> bdd_bundle(2 * n_bdds);
        for(size_t g = 0; g < L.bundles.size(); ++g)
            for(uint32_t q = 0; q < (32u >> L.bundles[g].logP); ++q)
            {
                const int32_t b = L.bundle_bdd[L.bundles[g].bdd_base + q];
                if(b >= 0) { bdd_bundle[2 * b] = (uint32_t)g; bdd_bundle[2 * b + 1] = q; }
            }

        d_bundles_.upload(L.bundles, stream_);
        d_chunks_.upload(L.chunks, stream_);
        d_desc_fwd_.upload(L.desc_fwd, stream_);
        d_desc_bwd_.upload(L.desc_bwd, stream_);
        d_desc_lane_.upload(L.desc_lane, stream_);
        d_hops_.upload(L.hops, stream_);
        d_topo_.upload(L.topo, stream_);
        d_bundle_bdd_.upload(L.bundle_bdd, stream_);
        d_bdd_bundle_.upload(bdd_bundle, stream_);
        d_ext2lay_.upload(L.ext2lay, stream_);
        d_ext_var_.upload(L.ext_var, stream_);
        d_ext_bdd_.upload(L.ext_bdd, stream_);
        d_bdd_ext_begin_.upload(L.bdd_ext_begin, stream_);
        d_var_lay_begin_.upload(L.var_lay_begin, stream_);
        d_var_lay_.upload(L.var_lay, stream_);
        d_sorted_ext_.upload(L.sorted_ext, stream_);
        d_nr_bdds_.upload(h_nr_bdds_per_var_, stream_);
        {   // per layer entry {variable, nr_bdds(variable)}: paired on the device (the host loop was a gather over tens of megabytes)
            DevBuf<int32_t> d_lay_var;
            d_lay_var.upload(L.lay_var, stream_);
            d_lay_vn_.alloc(L.n_lay);
            pair_lay_vn_kernel<<<blocks_for(L.n_lay), 256, 0, stream_>>>(d_lay_var.p, d_nr_bdds_.p, d_lay_vn_.p, (uint32_t)L.n_lay);
            CUDA_CHECK(cudaGetLastError());
            CUDA_CHECK(cudaStreamSynchronize(stream_));      // d_lay_var is freed here
        }
        {   // reciprocals for the lane-class kernels' normalisation (same rounding as the device division)
            std::vector<REAL> inv(INV_TAB);
            for(int i = 0; i < INV_TAB; ++i) inv[i] = (REAL)1 / (REAL)(i > 0 ? i : 1);
            d_inv_tab_.upload(inv, stream_);
            const int max_n = h_nr_bdds_per_var_.empty() ? 0 : *std::max_element(h_nr_bdds_per_var_.begin(), h_nr_bdds_per_var_.end());
            inv_count_ = (uint32_t)std::min<int>(INV_TAB, std::max(max_n, 0) + 1);
        }
        if(const char* e = std::getenv("BDDB200_NO_PDL")) pdl_ = std::atoi(e) == 0;

        d_cfr_.alloc(n_slots_); d_cft_.alloc(n_slots_);
        for(int i = 0; i < 2; ++i) { d_lohi_[i].alloc(2 * n_lay_); d_lohi_[i].zero(stream_); }
        d_mmd_.alloc(n_lay_); d_mmd_.zero(stream_);
        d_mm_lo_.alloc(n_lay_); d_mm_hi_.alloc(n_lay_);
        for(int i = 0; i < 3; ++i) { d_delta_[i].alloc(2 * n_vars_); d_delta_[i].zero(stream_); }
        d_delta_tmp_.alloc(2 * n_vars_);
        d_bdd_lb_.alloc(n_bdds_);
        d_lb_partial_.alloc(LB_BLOCKS + 1 + LB_SLOTS); d_lb_partial_.zero(stream_);
        CUDA_CHECK(cudaMallocHost(&h_lb_, LB_SLOTS * sizeof(double)));
        alloc_resident();

        configure_kernels();

        if(costs_hi != nullptr && n_costs > 0)
            update_costs_host(nullptr, 0, costs_hi, n_costs);
        CUDA_CHECK(cudaStreamSynchronize(stream_));   // host vectors of the layout go out of scope
    }

    // deep copy (the reference class is copyable: all members are thrust::device_vectors)
    bddb200_solver* clone() const override { return new SolverImpl(*this, 0); }
    SolverImpl(const SolverImpl& o, int)
    {
        precision = o.precision; device = o.device; deterministic_ = o.deterministic_;
        CUDA_CHECK(cudaSetDevice(device));
        CUDA_CHECK(cudaStreamSynchronize(o.stream_));
        CUDA_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking)); own_stream_ = true;
        n_sms_ = o.n_sms_; max_optin_ = o.max_optin_; warps_per_cta_ = o.warps_per_cta_; grid_small_ = o.grid_small_; forced_wpc_ = o.forced_wpc_;
        n_stages_ = o.n_stages_; n_stages_large_ = o.n_stages_large_;
        n_lane_shared_ = o.n_lane_shared_; layout_shared_vars_ = o.layout_shared_vars_;
        n_lane_ = o.n_lane_; lane_max_J_ = o.lane_max_J_; lane_max_hops_ = o.lane_max_hops_; lane_wpc_ = o.lane_wpc_; lane_grid_ = o.lane_grid_;
        lane_cls_first_ = o.lane_cls_first_; lane_cls_begin_ = o.lane_cls_begin_; lane_dense_ = o.lane_dense_;
        lane_chunk_hops_ = o.lane_chunk_hops_; lane_stages_ = o.lane_stages_; lane_stage_bytes_ = o.lane_stage_bytes_; lane_warp_smem_ = o.lane_warp_smem_;
        stage_small_ = o.stage_small_; stage_large_ = o.stage_large_; warp_smem_small_ = o.warp_smem_small_; warp_smem_large_ = o.warp_smem_large_;
        n_vars_ = o.n_vars_; n_bdds_ = o.n_bdds_; n_instr_ = o.n_instr_; n_ext_ = o.n_ext_; n_slots_ = o.n_slots_; n_lay_ = o.n_lay_;
        max_hops_ = o.max_hops_; n_bundles_ = o.n_bundles_; n_small_ = o.n_small_; tile_small_ = o.tile_small_; tile_large_ = o.tile_large_;
        h_nr_bdds_per_var_ = o.h_nr_bdds_per_var_; h_ext_var_ = o.h_ext_var_; h_ext_bdd_ = o.h_ext_bdd_;
        d_bundles_.clone_from(o.d_bundles_, stream_); d_chunks_.clone_from(o.d_chunks_, stream_);
        d_desc_fwd_.clone_from(o.d_desc_fwd_, stream_); d_desc_bwd_.clone_from(o.d_desc_bwd_, stream_); d_desc_lane_.clone_from(o.d_desc_lane_, stream_);
        d_hops_.clone_from(o.d_hops_, stream_); d_topo_.clone_from(o.d_topo_, stream_); d_bdd_bundle_.clone_from(o.d_bdd_bundle_, stream_);
        d_ext2lay_.clone_from(o.d_ext2lay_, stream_); d_bdd_ext_begin_.clone_from(o.d_bdd_ext_begin_, stream_);
        d_var_lay_begin_.clone_from(o.d_var_lay_begin_, stream_); d_var_lay_.clone_from(o.d_var_lay_, stream_); d_sorted_ext_.clone_from(o.d_sorted_ext_, stream_);
        d_lay_vn_.clone_from(o.d_lay_vn_, stream_); d_bundle_bdd_.clone_from(o.d_bundle_bdd_, stream_);
        d_ext_var_.clone_from(o.d_ext_var_, stream_); d_ext_bdd_.clone_from(o.d_ext_bdd_, stream_); d_nr_bdds_.clone_from(o.d_nr_bdds_, stream_);
        d_cfr_.clone_from(o.d_cfr_, stream_); d_cft_.clone_from(o.d_cft_, stream_);
        for(int i = 0; i < 2; ++i) d_lohi_[i].clone_from(o.d_lohi_[i], stream_);
        d_mmd_.clone_from(o.d_mmd_, stream_); d_mm_lo_.clone_from(o.d_mm_lo_, stream_); d_mm_hi_.clone_from(o.d_mm_hi_, stream_);
        for(int i = 0; i < 3; ++i) d_delta_[i].clone_from(o.d_delta_[i], stream_);
        d_inv_tab_.clone_from(o.d_inv_tab_, stream_); inv_count_ = o.inv_count_; pdl_ = o.pdl_;
        d_delta_tmp_.clone_from(o.d_delta_tmp_, stream_); d_bdd_lb_.clone_from(o.d_bdd_lb_, stream_); d_lb_partial_.clone_from(o.d_lb_partial_, stream_);
        CUDA_CHECK(cudaMallocHost(&h_lb_, LB_SLOTS * sizeof(double)));
        res_ok_ = o.res_ok_; res_grid_ = o.res_grid_; res_wpc_ = o.res_wpc_; res_warp_smem_ = o.res_warp_smem_;
        res_vars_per_bundle_ = o.res_vars_per_bundle_; res_own_cap_ = o.res_own_cap_; res_own_off_ = o.res_own_off_;
        d_contrib_.clone_from(o.d_contrib_, stream_); d_sums_.clone_from(o.d_sums_, stream_); d_lb_part_.clone_from(o.d_lb_part_, stream_);
        if(res_ok_) CUDA_CHECK(cudaMallocHost(&h_lb_part_, std::max<size_t>(n_lane_, 1) * sizeof(double)));
        res_pass_ = o.res_pass_; exch_in_contrib_ = o.exch_in_contrib_; lb_from_resident_ = o.lb_from_resident_;
        cc_ = o.cc_; dcur_ = o.dcur_; delta_needs_norm_ = o.delta_needs_norm_;
        forward_valid_ = o.forward_valid_; backward_valid_ = o.backward_valid_; lb_valid_ = o.lb_valid_; lb_ = o.lb_;
        configure_kernels();
        CUDA_CHECK(cudaStreamSynchronize(stream_));
    }

    // ---- serialisation (the reference class is cereal-serialisable, bdd_cuda_base.cu:1486-1544 / bdd_cuda_parallel_mma.cu:475-490: all device
    // vectors and counters; its pybind module pickles solvers that way, bdd_cuda_parallel_mma_py.cu:29-38).  One blob: header, then every
    // member the copy constructor copies, device buffers as raw bytes.  The launch plan is part of it, so a blob loads on the same GPU model.
    struct SaveArchive {
        std::vector<unsigned char>& out; cudaStream_t st;
        void raw(const void* p, size_t n) { const unsigned char* b = static_cast<const unsigned char*>(p); out.insert(out.end(), b, b + n); }
        template<typename T> void pod(T& v) { raw(&v, sizeof(T)); }
        template<typename T> void vec(std::vector<T>& v) { uint64_t n = v.size(); pod(n); if(n) raw(v.data(), n * sizeof(T)); }
        template<typename T> void dev(DevBuf<T>& b)
        {
            uint64_t n = b.n; pod(n);
            const size_t off = out.size();
            out.resize(off + n * sizeof(T));
            if(n) CUDA_CHECK(cudaMemcpyAsync(out.data() + off, b.p, n * sizeof(T), cudaMemcpyDeviceToHost, st));
            CUDA_CHECK(cudaStreamSynchronize(st));        // `out` may reallocate on the next append
        }
    };
    struct LoadArchive {
        const unsigned char* p; size_t left; cudaStream_t st;
        void raw(void* dst, size_t n)
        {
            if(n > left) throw api_error(BDDB200_ERR_INVALID_ARGUMENT, "bddb200_load: truncated blob");
            std::memcpy(dst, p, n); p += n; left -= n;
        }
        template<typename T> void pod(T& v) { raw(&v, sizeof(T)); }
        template<typename T> void vec(std::vector<T>& v)
        {
            uint64_t n = 0; pod(n);
            if(n * sizeof(T) > left) throw api_error(BDDB200_ERR_INVALID_ARGUMENT, "bddb200_load: truncated blob");
            v.resize(n); if(n) raw(v.data(), n * sizeof(T));
        }
        template<typename T> void dev(DevBuf<T>& b)
        {
            uint64_t n = 0; pod(n);
            if(n * sizeof(T) > left) throw api_error(BDDB200_ERR_INVALID_ARGUMENT, "bddb200_load: truncated blob");
            b.alloc(n);
            if(n) CUDA_CHECK(cudaMemcpyAsync(b.p, p, n * sizeof(T), cudaMemcpyHostToDevice, st));
            p += n * sizeof(T); left -= n * sizeof(T);
        }
    };
    template<typename A>
    void visit_state(A& ar)
    {
        ar.pod(deterministic_);
        ar.pod(n_sms_); ar.pod(max_optin_); ar.pod(warps_per_cta_); ar.pod(grid_small_); ar.pod(forced_wpc_); ar.pod(n_stages_); ar.pod(n_stages_large_);
        ar.pod(n_lane_shared_); ar.pod(layout_shared_vars_);
        ar.pod(n_lane_); ar.pod(lane_max_J_); ar.pod(lane_max_hops_); ar.pod(lane_wpc_); ar.pod(lane_grid_); ar.vec(lane_cls_first_); ar.vec(lane_cls_begin_); ar.pod(lane_dense_);
        ar.pod(lane_chunk_hops_); ar.pod(lane_stages_); ar.pod(lane_stage_bytes_); ar.pod(lane_warp_smem_);
        ar.pod(stage_small_); ar.pod(stage_large_); ar.pod(warp_smem_small_); ar.pod(warp_smem_large_);
        ar.pod(n_vars_); ar.pod(n_bdds_); ar.pod(n_instr_); ar.pod(n_ext_); ar.pod(n_slots_); ar.pod(n_lay_);
        ar.pod(max_hops_); ar.pod(n_bundles_); ar.pod(n_small_); ar.pod(tile_small_); ar.pod(tile_large_);
        ar.vec(h_nr_bdds_per_var_); ar.vec(h_ext_var_); ar.vec(h_ext_bdd_);
        ar.dev(d_bundles_); ar.dev(d_chunks_); ar.dev(d_desc_fwd_); ar.dev(d_desc_bwd_); ar.dev(d_desc_lane_);
        ar.dev(d_hops_); ar.dev(d_topo_); ar.dev(d_bdd_bundle_); ar.dev(d_ext2lay_); ar.dev(d_bdd_ext_begin_);
        ar.dev(d_var_lay_begin_); ar.dev(d_var_lay_); ar.dev(d_sorted_ext_); ar.dev(d_lay_vn_); ar.dev(d_bundle_bdd_);
        ar.dev(d_ext_var_); ar.dev(d_ext_bdd_); ar.dev(d_nr_bdds_); ar.dev(d_cfr_); ar.dev(d_cft_);
        ar.dev(d_lohi_[0]); ar.dev(d_lohi_[1]); ar.dev(d_mmd_); ar.dev(d_mm_lo_); ar.dev(d_mm_hi_);
        ar.dev(d_delta_[0]); ar.dev(d_delta_[1]); ar.dev(d_delta_[2]);
        ar.dev(d_inv_tab_); ar.pod(inv_count_); ar.pod(pdl_);
        ar.dev(d_delta_tmp_); ar.dev(d_bdd_lb_); ar.dev(d_lb_partial_);
        ar.pod(res_ok_); ar.pod(res_grid_); ar.pod(res_wpc_); ar.pod(res_warp_smem_); ar.pod(res_vars_per_bundle_); ar.pod(res_own_cap_); ar.pod(res_own_off_);
        ar.dev(d_contrib_); ar.dev(d_sums_); ar.dev(d_lb_part_);
        ar.pod(res_pass_); ar.pod(exch_in_contrib_); ar.pod(lb_from_resident_);
        ar.pod(cc_); ar.pod(dcur_); ar.pod(delta_needs_norm_); ar.pod(forward_valid_); ar.pod(backward_valid_); ar.pod(lb_valid_); ar.pod(lb_); ar.pod(lb_sum_clean_);
    }
    static constexpr uint64_t SAVE_MAGIC = 0x3130533030324442ull;      // "BD200S01"
    void save(std::vector<unsigned char>& out) const override
    {
        if(ext_delta_[0] != nullptr || xc_.mode != 0)
            throw api_error(BDDB200_ERR_STATE, "save: a solver whose sum buffers live in caller-owned (symmetric) memory cannot be serialised");
        SolverImpl& self = const_cast<SolverImpl&>(*this);       // the visitor only reads here
        self.set_device();
        CUDA_CHECK(cudaStreamSynchronize(stream_));
        SaveArchive ar{out, stream_};
        uint64_t magic = SAVE_MAGIC; int32_t prec = precision;
        ar.pod(magic); ar.pod(prec);
        self.visit_state(ar);
    }
    // the loading constructor: `blob` points behind the header
    SolverImpl(const unsigned char* blob, size_t bytes, int dev)
    {
        precision = sizeof(REAL) == 8 ? BDDB200_DOUBLE : BDDB200_FLOAT;
        device = dev;
        int count = 0;
        if(cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
            throw api_error(BDDB200_ERR_NO_DEVICE, "no CUDA device available: libbdd_b200 has no CPU fallback");
        if(device < 0 || device >= count) throw api_error(BDDB200_ERR_INVALID_ARGUMENT, "invalid device ordinal");
        CUDA_CHECK(cudaSetDevice(device));
        CUDA_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking)); own_stream_ = true;
        LoadArchive ar{blob, bytes, stream_};
        visit_state(ar);
        int sms = 0, optin = 0;
        CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        CUDA_CHECK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        if(sms != n_sms_ || optin != (int)max_optin_)
            throw api_error(BDDB200_ERR_INVALID_ARGUMENT, "bddb200_load: the blob was saved on a different GPU model (its launch plan does not apply here)");
        CUDA_CHECK(cudaMallocHost(&h_lb_, LB_SLOTS * sizeof(double)));
        if(res_ok_) CUDA_CHECK(cudaMallocHost(&h_lb_part_, std::max<size_t>(n_lane_, 1) * sizeof(double)));
        configure_kernels();
        CUDA_CHECK(cudaStreamSynchronize(stream_));
    }

    ~SolverImpl() override
    {
        cudaSetDevice(device);
        if(xc_.mode == 4 && push_pending_)
        {   // peers may still be adding to this rank's (caller-owned, symmetric) sum buffers: they must be done before the caller frees them
            try { push_barrier(1); cudaStreamSynchronize(stream_); } catch(...) {}
        }
        if(graph_exec_) cudaGraphExecDestroy(graph_exec_);
        for(StepGraph& c : step_graphs_) cudaGraphExecDestroy(c.exec);
        if(h_lb_) cudaFreeHost(h_lb_);
        if(h_lb_part_) cudaFreeHost(h_lb_part_);
        if(h_round_counts_) cudaFreeHost(h_round_counts_);
        if(h_stage_) { cudaFreeHost(h_stage_); cudaEventDestroy(stage_free_); }
        if(own_stream_ && stream_) cudaStreamDestroy(stream_);
    }

    size_t nr_variables() const override { return n_vars_; }
    size_t nr_bdds() const override { return n_bdds_; }
    size_t nr_layers() const override { return n_ext_; }
    size_t nr_bdd_nodes() const override { return n_instr_; }
    size_t nr_hops() const override { return max_hops_ > 0 ? max_hops_ - 1 : 0; }
    void nr_bdds_per_var(int32_t* out) const override { std::memcpy(out, h_nr_bdds_per_var_.data(), sizeof(int32_t) * n_vars_); }
    void layer_primal_indices(int32_t* out) const override { std::memcpy(out, h_ext_var_.data(), sizeof(int32_t) * n_ext_); }
    void layer_bdd_indices(int32_t* out) const override { std::memcpy(out, h_ext_bdd_.data(), sizeof(int32_t) * n_ext_); }

    // ---------------------------------------------------------------- sweep launches ---
    template<int MODE, bool FORWARD>
    void launch_sweep(SweepArgs<REAL> a)
    {
        auto kern = sweep_kernel<REAL, MODE, FORWARD>;
        // every backward sweep ADDS the roots' values to lb_sum; every forward sweep clears it (block 0).  A backward sweep that
        // does not directly follow a forward sweep (forward_mm, lower_bound(), backward_mm: the bound in between runs a plain
        // backward sweep) must clear it first.
        if(MODE == MODE_MMA) ensure_sums();
        if(!FORWARD && !deterministic_ && !lb_sum_clean_) zero_lb_sum();
        lb_sum_clean_ = FORWARD;
        if(!FORWARD) lb_from_resident_ = false;
        if(n_lane_ > 0)
        {   // lane-local class: bundles [0, n_lane)
            a.desc = reinterpret_cast<const uint32_t*>(d_desc_lane_.p);
            a.bundle_first = 0; a.bundle_count = (uint32_t)n_lane_; a.tile_slots = 0;
            a.stage_bytes = lane_stage_bytes_; a.n_stages = lane_stages_; a.chunk_hops = lane_chunk_hops_; a.warp_smem_bytes = lane_warp_smem_;
            a.inv_tab_g = d_inv_tab_.p; a.inv_count = inv_count_;
            a.bundles_per_cta = (uint32_t)(n_lane_ / lane_grid_); a.bundles_rem = (uint32_t)(n_lane_ % lane_grid_);
            a.zero_pairs_per_bundle = (uint32_t)((n_vars_ + n_lane_ - 1) / n_lane_);
            a.n_classes = (uint32_t)lane_cls_begin_.size();
            for(size_t c = 0; c < lane_cls_begin_.size(); ++c) { a.cls_first[c] = lane_cls_first_[c]; a.cls_begin[c] = lane_cls_begin_[c]; }
            // programmatic dependent launch: the kernel's start-up (descriptor, static topology, variable indices) overlaps the
            // tail of the previous kernel in the stream; it waits (griddepcontrol.wait) before touching anything a pass writes
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(lane_grid_); cfg.blockDim = dim3(lane_wpc_ * 32);
            cfg.dynamicSmemBytes = (size_t)lane_wpc_ * lane_warp_smem_; cfg.stream = stream_;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr; cfg.numAttrs = pdl_ ? 1 : 0;
            if(MODE == MODE_MMA && a.push_counters != nullptr)
            {   // multi-GPU push exchange: the pass adds the shared variables' differences to every rank's buffer itself
                if(lane_dense_) CUDA_CHECK(cudaLaunchKernelEx(&cfg, sweep_lane_kernel<REAL, MODE_MMA, FORWARD, false, 768, true>, a));
                else CUDA_CHECK(cudaLaunchKernelEx(&cfg, sweep_lane_kernel<REAL, MODE_MMA, FORWARD, false, BDDB200_LANE_MAX_THREADS, true>, a));
            }
            else if(MODE == MODE_MMA && !deterministic_ && lane_dense_) CUDA_CHECK(cudaLaunchKernelEx(&cfg, sweep_lane_kernel<REAL, MODE_MMA, FORWARD, false, 768>, a));
            else if(MODE == MODE_MMA && deterministic_) CUDA_CHECK(cudaLaunchKernelEx(&cfg, sweep_lane_kernel<REAL, MODE, FORWARD, MODE == MODE_MMA>, a));
            else CUDA_CHECK(cudaLaunchKernelEx(&cfg, sweep_lane_kernel<REAL, MODE, FORWARD, false>, a));
            ++launches_;
            a.zero_buf = nullptr;
        }
        a.desc = FORWARD ? d_desc_fwd_.p : d_desc_bwd_.p;
        if(n_small_ > 0)
        {
            a.bundle_first = (uint32_t)n_lane_; a.bundle_count = (uint32_t)n_small_; a.tile_slots = tile_small_;
            a.stage_bytes = stage_small_; a.n_stages = n_stages_; a.warp_smem_bytes = warp_smem_small_;
            kern<<<grid_small_, warps_per_cta_ * 32, INV_TAB_BYTES + (size_t)warps_per_cta_ * warp_smem_small_, stream_>>>(a);
            ++launches_;
        }
        if(n_bundles_ > n_lane_ + n_small_)
        {
            a.bundle_first = (uint32_t)(n_lane_ + n_small_); a.bundle_count = (uint32_t)(n_bundles_ - n_lane_ - n_small_); a.tile_slots = tile_large_;
            a.stage_bytes = stage_large_; a.n_stages = n_stages_large_; a.warp_smem_bytes = warp_smem_large_;
            if(n_small_ > 0) a.zero_buf = nullptr;
            kern<<<(unsigned)(n_bundles_ - n_lane_ - n_small_), 32, INV_TAB_BYTES + warp_smem_large_, stream_>>>(a);
            ++launches_;
        }
        CUDA_CHECK(cudaGetLastError());
    }

    // Shared memory of one warp: n_stages pipeline stages, two frontier buffers, the mbarriers.
    static uint32_t warp_smem(uint32_t n_stages, uint32_t stage_bytes, uint32_t tile_slots)
    {
        return (uint32_t)((((size_t)n_stages * stage_bytes + 2 * (size_t)tile_slots * sizeof(REAL) + 8 * (size_t)n_stages) + 127) & ~(size_t)127);
    }

    // Choose warps per CTA for the small class (bundles that fit the stage budget) and the
    // pipeline depth of the large class (one warp per CTA).
    void plan_launch()
    {
        const size_t budget = (size_t)max_optin_ - INV_TAB_BYTES;
        warp_smem_small_ = warp_smem(n_stages_, stage_small_, tile_small_);
        if(n_small_ > 0 && warp_smem_small_ > budget)
            throw api_error(BDDB200_ERR_TOO_WIDE, "stage_bytes * n_stages does not fit the shared memory of one SM");
        const unsigned max_wps = (unsigned)std::max<size_t>(1, std::min<size_t>(16, budget / std::max<uint32_t>(warp_smem_small_, 1u)));
        // the kernel deals the bundles evenly over the CTAs (each gets floor or ceil of n / grid)
        if(forced_wpc_) { warps_per_cta_ = std::min(forced_wpc_, max_wps); grid_small_ = blocks_for(n_small_, warps_per_cta_); }
        else if(n_small_ <= (size_t)n_sms_ * max_wps)
        {   // less than one wave: one CTA on every SM
            grid_small_ = (unsigned)std::max<size_t>(1, std::min<size_t>(n_small_, (size_t)n_sms_));
            warps_per_cta_ = (unsigned)std::max<size_t>(1, (n_small_ + grid_small_ - 1) / grid_small_);
        }
        else { warps_per_cta_ = std::min(4u, max_wps); grid_small_ = blocks_for(n_small_, warps_per_cta_); }
        if(n_lane_ > 0) plan_lane_launch();
        if(n_bundles_ > n_lane_ + n_small_)
        {
            n_stages_large_ = n_stages_;
            while(n_stages_large_ > 2 && warp_smem(n_stages_large_, stage_large_, tile_large_) > budget) --n_stages_large_;
            warp_smem_large_ = warp_smem(n_stages_large_, stage_large_, tile_large_);
            if(warp_smem_large_ > budget)
                throw api_error(BDDB200_ERR_TOO_WIDE, "widest BDD layer needs " + std::to_string(warp_smem_large_) +
                                " bytes of shared memory per warp (limit " + std::to_string(budget) + "); split the BDD (split_qbdd)");
        }
    }

    // Lane-local class: warps per CTA, CTAs per SM, pipeline depth and hops per stage.
    void plan_lane_launch()
    {
        auto env_u = [](const char* name, unsigned dflt) { const char* e = std::getenv(name); return e ? (unsigned)std::atoi(e) : dflt; };
        const size_t per_hop = lane_hop_bytes(lane_max_J_, sizeof(REAL));
        lane_stages_ = std::min((unsigned)LANE_MAX_STAGES, std::max(1u, env_u("BDDB200_LANE_STAGES", 2)));
        unsigned ctas_per_sm = 1;
        const unsigned forced = env_u("BDDB200_LANE_WARPS_PER_CTA", 0);
        if(forced) { lane_wpc_ = std::min(forced, 16u); lane_grid_ = blocks_for(n_lane_, lane_wpc_); ctas_per_sm = std::max(1u, env_u("BDDB200_LANE_CTAS_PER_SM", 1)); }
        else if(n_lane_ <= (size_t)n_sms_ * 16)
        {   // at most one wave of 16 warps per SM: one CTA on every SM, the bundles dealt evenly (no second, partly filled wave);
            // the fewer warps, the deeper the stages the shared memory is spent on
            lane_grid_ = (unsigned)std::min<size_t>(n_lane_, (size_t)n_sms_);
            lane_wpc_ = (unsigned)((n_lane_ + lane_grid_ - 1) / lane_grid_);
        }
        else
        {   // several waves: between 8 and 16 resident warps per SM, as many as leave every stage about six hops (the per-chunk cost --
            // four bulk-copy issues, one batch of gathers -- is about two hops of arithmetic), adjusted so that the last wave is
            // as full as possible (time ~ waves x warps per SM)
            const size_t w_max = std::min<size_t>(16, std::max<size_t>(8, (size_t)(227 * 1024) / (lane_stages_ * 6 * per_hop)));
            size_t best_w = w_max, best_cost = ~(size_t)0;
            for(size_t w = w_max; w >= 8; --w)
            {
                const size_t waves = (n_lane_ + (size_t)n_sms_ * w - 1) / ((size_t)n_sms_ * w);
                if(waves * w < best_cost) { best_cost = waves * w; best_w = w; }
            }
            // float, many waves: the register-capped build of the MMA pass runs 24 warps per SM (3 CTAs of 8) if that leaves >= 3 hops per stage
            lane_dense_ = sizeof(REAL) == 4 && !deterministic_ && env_u("BDDB200_NO_DENSE", 0) == 0 && n_lane_ >= (size_t)n_sms_ * 48
                          && ((size_t)(75 * 1024) - INV_TAB * sizeof(REAL) - 1024) / 8 / (lane_stages_ * per_hop) >= 3;
            if(lane_dense_) { ctas_per_sm = 3; lane_wpc_ = 8; }
            else if(best_w >= 12 && best_w % 2 == 0) { ctas_per_sm = 2; lane_wpc_ = (unsigned)(best_w / 2); } else lane_wpc_ = (unsigned)best_w;
            lane_grid_ = blocks_for(n_lane_, lane_wpc_);
        }
        const size_t sm_total = 228 * 1024;                             // per SM; every resident CTA reserves 1 KiB
        const size_t static_smem = INV_TAB * sizeof(REAL) + 16 * LANE_MAX_STAGES * 8 + 128;   // reciprocal table + mbarriers
        size_t cta_budget = std::min<size_t>((size_t)max_optin_, sm_total / ctas_per_sm - 1024) - static_smem;
        const size_t warp_budget = cta_budget / lane_wpc_;
        size_t hops = (warp_budget - 128) / (lane_stages_ * per_hop);
        if(hops < 1) throw api_error(BDDB200_ERR_TOO_WIDE, "lane-class stage does not fit the shared memory of one SM");
        hops = std::min<size_t>(hops, lane_max_hops_);
        if(const unsigned f = env_u("BDDB200_LANE_CHUNK_HOPS", 0)) hops = std::min<size_t>(hops, f);
        size_t nc = (lane_max_hops_ + hops - 1) / hops;
        if(lane_stages_ == 1 && nc > 1)
        {   // a single stage only works when a whole bundle fits into it
            lane_stages_ = 2;
            hops = std::max<size_t>(1, (warp_budget - 128) / (lane_stages_ * per_hop));
            hops = std::min<size_t>(hops, lane_max_hops_);
            nc = (lane_max_hops_ + hops - 1) / hops;
        }
        hops = (lane_max_hops_ + nc - 1) / nc;                         // chunks of equal length
        lane_chunk_hops_ = (uint32_t)hops;
        lane_stage_bytes_ = (uint32_t)(hops * per_hop);
        lane_warp_smem_ = (uint32_t)((((size_t)lane_stages_ * lane_stage_bytes_) + 127) & ~(size_t)127);
    }

    // On-chip form (resident.cuh): every bundle is lane class, the collection is at most one wave of 16 warps per SM and the
    // state of the bundles of one CTA fits the shared memory of an SM.
    void plan_resident()
    {
        res_ok_ = false;
        // Opt-in (BDDB200_RESIDENT=1): on B200 the exchange of the per-variable sums bounds a pass of such collections -- two scattered
        // L2 accesses per layer entry, whichever way they are made -- and the streaming kernels are already at that bound
        // (profiles/r02_*), so keeping the state on chip buys nothing there; the kernel is kept for its bit-reproducible sums.
        if(std::getenv("BDDB200_RESIDENT") == nullptr || std::atoi(std::getenv("BDDB200_RESIDENT")) == 0) return;
        if(n_lane_ == 0 || n_lane_ != n_bundles_ || n_lane_ > (size_t)n_sms_ * 16 || n_vars_ >= ((size_t)1 << 26)) return;
        res_grid_ = (unsigned)std::min<size_t>(n_lane_, (size_t)n_sms_);
        res_wpc_ = (unsigned)((n_lane_ + res_grid_ - 1) / res_grid_);
        const uint32_t state_bytes = (uint32_t)((resident_bundle_bytes(lane_max_hops_, lane_max_J_, sizeof(REAL)) + 127) & ~(size_t)127);
        const size_t static_smem = 16 * 8 + 256;
        if((size_t)res_wpc_ * state_bytes + static_smem > (size_t)max_optin_) return;
        // owner duty: bundle g owns vars_per_bundle consecutive variables; their layer lists are staged in shared memory when they fit
        res_vars_per_bundle_ = (uint32_t)((n_vars_ + n_lane_ - 1) / n_lane_);
        size_t max_list = 0;
        for(size_t g = 0; g < n_lane_; ++g)
        {
            const size_t b = std::min(n_vars_, g * (size_t)res_vars_per_bundle_), e = std::min(n_vars_, b + res_vars_per_bundle_);
            max_list = std::max<size_t>(max_list, L_var_lay_begin_[e] - L_var_lay_begin_[b]);
        }
        const size_t own_bytes = ((((size_t)res_vars_per_bundle_ + 1 + 3) & ~(size_t)3) + max_list) * 4;
        res_own_off_ = state_bytes; res_own_cap_ = 0; res_warp_smem_ = state_bytes;
        if((size_t)res_wpc_ * (state_bytes + ((own_bytes + 127) & ~(size_t)127)) + static_smem <= (size_t)max_optin_)
        {
            res_own_cap_ = (uint32_t)std::max<size_t>(max_list, 1);
            res_warp_smem_ = (uint32_t)(state_bytes + ((own_bytes + 127) & ~(size_t)127));
        }
        int coop = 0;
        CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
        res_ok_ = coop != 0;
    }
    void alloc_resident()
    {
        if(!res_ok_) return;
        d_contrib_.alloc(n_lay_ * ExchangeRec<REAL>::CONTRIB); d_contrib_.zero(stream_);
        d_sums_.alloc((n_vars_ + 1) * ExchangeRec<REAL>::SUM); d_sums_.zero(stream_);
        d_lb_part_.alloc(n_lane_); d_lb_part_.zero(stream_);
        CUDA_CHECK(cudaMallocHost(&h_lb_part_, std::max<size_t>(n_lane_, 1) * sizeof(double)));
    }
    bool use_resident() const { return res_ok_ && delta_in_override_ == nullptr && ext_delta_[0] == nullptr; }

    // After a launch of the on-chip kernel the pending per-variable sums exist as the contribution records of its last pass (and
    // as deffered_mm_diff_): put them into the rotating sum buffer the streaming kernels and the accessors read.
    void ensure_sums()
    {
        if(!exch_in_contrib_) return;
        delta_segsum_kernel<REAL><<<blocks_for(n_vars_), 256, 0, stream_>>>(d_var_lay_begin_.p, d_var_lay_.p, d_mmd_.p, dbuf(dcur_), (uint32_t)n_vars_);
        ++launches_;
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaMemsetAsync(dbuf((dcur_ + 1) % 3), 0, sizeof(REAL) * 2 * n_vars_, stream_));      // the next pass accumulates into it
        delta_needs_norm_ = true;
        exch_in_contrib_ = false;
    }

    // n iterations in ONE cooperative launch; cost_from_terminal is recomputed on chip first when it is stale
    void launch_resident(double omega, size_t n, unsigned long long* trace = nullptr)
    {
        using R2 = typename real2<REAL>::type;
        while(n > 0)
        {
            const size_t k = std::min<size_t>(n, 1u << 16);
            uint32_t published = 0;
            if(!exch_in_contrib_)
            {   // hand the pending sums of the rotating buffer over to the on-chip kernel's exchange
                ++res_pass_;
                publish_sums_kernel<REAL><<<blocks_for(n_vars_), 256, 0, stream_>>>(dbuf(dcur_), d_nr_bdds_.p, d_sums_.p, res_pass_, delta_needs_norm_ ? 1 : 0, (uint32_t)n_vars_);
                ++launches_;
                CUDA_CHECK(cudaGetLastError());
                published = 1;
            }
            ResidentArgs<REAL> a{};
            a.desc = d_desc_lane_.p; a.topo = d_topo_.p; a.lay_vn = d_lay_vn_.p; a.bundle_bdd = d_bundle_bdd_.p;
            a.cfr = d_cfr_.p; a.cft = d_cft_.p; a.lohi = reinterpret_cast<R2*>(d_lohi_[cc_].p); a.mmd = d_mmd_.p;
            a.contrib = d_contrib_.p; a.sums = d_sums_.p; a.var_lay_begin = d_var_lay_begin_.p; a.var_lay = d_var_lay_.p; a.nr_bdds = d_nr_bdds_.p;
            a.n_vars = (uint32_t)n_vars_; a.vars_per_bundle = res_vars_per_bundle_; a.own_list_cap = res_own_cap_; a.own_smem_off = res_own_off_;
            a.pass0 = res_pass_; a.sums_published = published;
            a.bdd_lb = d_bdd_lb_.p; a.lb_part = d_lb_part_.p;
            a.omega = (REAL)omega; a.n_iterations = (uint32_t)k; a.init_backward = backward_valid_ ? 0u : 1u;
            a.n_bundles = (uint32_t)n_lane_; a.bundles_per_cta = (uint32_t)(n_lane_ / res_grid_); a.bundles_rem = (uint32_t)(n_lane_ % res_grid_);
            a.warp_smem_bytes = res_warp_smem_;
            a.n_classes = (uint32_t)lane_cls_begin_.size();
            for(size_t c = 0; c < lane_cls_begin_.size(); ++c) { a.cls_first[c] = lane_cls_first_[c]; a.cls_begin[c] = lane_cls_begin_[c]; }
            a.trace = trace;
            void* params[] = { &a };
            const void* kern = res_wpc_ <= 8 ? (const void*)resident_kernel<REAL, 256> : (const void*)resident_kernel<REAL, 512>;
            CUDA_CHECK(cudaLaunchCooperativeKernel(kern, dim3(res_grid_), dim3(res_wpc_ * 32), params, (size_t)res_wpc_ * res_warp_smem_, stream_));
            ++launches_;
            res_pass_ += (uint32_t)(2 * k);
            exch_in_contrib_ = true; delta_needs_norm_ = true;
            forward_valid_ = false; backward_valid_ = true; lb_valid_ = false; lb_sum_clean_ = false; lb_from_resident_ = true;
            n -= k;
        }
    }

    void configure_kernels()
    {
        if(res_ok_)
        {
            CUDA_CHECK(cudaFuncSetAttribute(resident_kernel<REAL, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)res_wpc_ * res_warp_smem_)));
            CUDA_CHECK(cudaFuncSetAttribute(resident_kernel<REAL, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)res_wpc_ * res_warp_smem_)));
        }
        const int need_lane = (int)((size_t)lane_wpc_ * lane_warp_smem_);
        if(n_lane_ > 0 && need_lane > 40 * 1024)
        {
            CUDA_CHECK(cudaFuncSetAttribute(sweep_lane_kernel<REAL, MODE_MMA, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, need_lane));
            CUDA_CHECK(cudaFuncSetAttribute(sweep_lane_kernel<REAL, MODE_MMA, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, need_lane));
            CUDA_CHECK(cudaFuncSetAttribute(sweep_lane_kernel<REAL, MODE_MMA, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, need_lane));
            CUDA_CHECK(cudaFuncSetAttribute(sweep_lane_kernel<REAL, MODE_MMA, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, need_lane));
            CUDA_CHECK(cudaFuncSetAttribute(sweep_lane_kernel<REAL, MODE_PLAIN, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, need_lane));
            CUDA_CHECK(cudaFuncSetAttribute(sweep_lane_kernel<REAL, MODE_PLAIN, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, need_lane));
            CUDA_CHECK(cudaFuncSetAttribute(sweep_lane_kernel<REAL, MODE_MM, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, need_lane));
            CUDA_CHECK(cudaFuncSetAttribute(sweep_lane_kernel<REAL, MODE_MMA, true, false, 768>, cudaFuncAttributeMaxDynamicSharedMemorySize, need_lane));
            CUDA_CHECK(cudaFuncSetAttribute(sweep_lane_kernel<REAL, MODE_MMA, false, false, 768>, cudaFuncAttributeMaxDynamicSharedMemorySize, need_lane));
            CUDA_CHECK(cudaFuncSetAttribute(sweep_lane_kernel<REAL, MODE_MMA, true, false, 768, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, need_lane));
            CUDA_CHECK(cudaFuncSetAttribute(sweep_lane_kernel<REAL, MODE_MMA, false, false, 768, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, need_lane));
            CUDA_CHECK(cudaFuncSetAttribute(sweep_lane_kernel<REAL, MODE_MMA, true, false, BDDB200_LANE_MAX_THREADS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, need_lane));
            CUDA_CHECK(cudaFuncSetAttribute(sweep_lane_kernel<REAL, MODE_MMA, false, false, BDDB200_LANE_MAX_THREADS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, need_lane));
        }
        const int need = (int)(INV_TAB_BYTES + std::max<size_t>((size_t)warps_per_cta_ * warp_smem_small_, warp_smem_large_));
        if(need > 48 * 1024)
        {
            CUDA_CHECK(cudaFuncSetAttribute(sweep_kernel<REAL, MODE_MMA, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, need));
            CUDA_CHECK(cudaFuncSetAttribute(sweep_kernel<REAL, MODE_MMA, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, need));
            CUDA_CHECK(cudaFuncSetAttribute(sweep_kernel<REAL, MODE_PLAIN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, need));
            CUDA_CHECK(cudaFuncSetAttribute(sweep_kernel<REAL, MODE_PLAIN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, need));
            CUDA_CHECK(cudaFuncSetAttribute(sweep_kernel<REAL, MODE_MM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, need));
        }
    }

    void zero_lb_sum()
    {
        if(!deterministic_) CUDA_CHECK(cudaMemsetAsync(d_lb_partial_.p + LB_BLOCKS + 1, 0, LB_SLOTS * sizeof(double), stream_));
    }

    SweepArgs<REAL> base_args() const
    {
        SweepArgs<REAL> a{};
        using R2 = typename real2<REAL>::type;
        a.chunks = d_chunks_.p; a.topo = d_topo_.p; a.lay_vn = d_lay_vn_.p; a.bundle_bdd = d_bundle_bdd_.p;
        a.cfr = d_cfr_.p; a.cft = d_cft_.p;
        a.lohi_in = reinterpret_cast<const R2*>(d_lohi_[cc_].p); a.lohi_out = reinterpret_cast<R2*>(d_lohi_[cc_ ^ 1].p);
        a.mmd = d_mmd_.p; a.mm_lo_out = d_mm_lo_.p; a.mm_hi_out = d_mm_hi_.p; a.bdd_lb = d_bdd_lb_.p;
        a.omega = 0; a.n_zero = (uint32_t)(2 * n_vars_); a.trace = trace_;
        a.lb_sum = deterministic_ ? nullptr : d_lb_partial_.p + LB_BLOCKS + 1;
        return a;
    }

    // one MMA pass on explicit delta buffers.  zero_buf is cleared by the same launch.
    template<bool FORWARD>
    void mma_pass(double omega, const REAL* delta_in, bool normalize_in, REAL* delta_out, REAL* zero_buf)
    {
        SweepArgs<REAL> a = base_args();
        a.omega = (REAL)omega;
        a.delta_in = delta_in; a.delta_out = delta_out; a.zero_buf = zero_buf;
        const bool own_sums = delta_in_override_ != nullptr && delta_in == dbuf(dcur_);     // not for forward_mm / backward_mm on a caller's vector
        a.delta_in_shared = own_sums ? delta_in_override_ : delta_in; a.n_shared_vars = own_sums ? (uint32_t)n_shared_vars_ : 0u;
        if(xc_.mode == 4 && delta_in == dbuf(dcur_) && delta_out == dbuf((dcur_ + 1) % 3))
        {   // push exchange: shared variables' differences go to every rank's buffer through the multicast mapping; the flag barrier
            // is the tail of this launch and the prologue of the next one
            a.push_peers = reinterpret_cast<REAL* const*>(const_cast<void* const*>(xc_.peers)); a.push_offset = (size_t)(delta_out - dbuf(0)); a.push_mask = d_push_mask_.p;
            a.delta_out_mc = xc_.mc_in != nullptr ? const_cast<REAL*>(xc_.mc_in) + a.push_offset : nullptr;
            a.push_debug = push_debug_;
            a.n_push_vars = (uint32_t)(xc_.n_exchange / 2); a.push_n_shared_bundles = (uint32_t)push_n_shared_bundles_;
            a.push_counters = d_xc_counters_.p; a.push_flags = xc_.flags; a.push_my_flags = xc_.my_flags;
            a.push_world = xc_.world; a.push_rank = xc_.rank;
            // this pass ends push barrier number e (counted mod 3: a rank is never more than one barrier ahead of a peer)
            const uint32_t e = (xc_phase_ + 1u) % 3u;
            a.push_send_phase = e; a.push_stale_phase = (e + 1u) % 3u;       // stale = the barrier before the previous one
            xc_phase_ = e;
            push_pending_ = true;
        }
        a.normalize_in = !normalize_in ? NORM_NONE : (deterministic_ ? NORM_DIVIDE : NORM_RECIPROCAL);
        a.accumulate = deterministic_ ? 0 : 1;
        launch_sweep<MODE_MMA, FORWARD>(a);
        cc_ ^= 1;   // thrust::swap(lo_cost_, lo_cost_out_), bdd_cuda_parallel_mma.cu:246-247
        if(deterministic_)
        {
            delta_segsum_kernel<REAL><<<blocks_for(n_vars_), 256, 0, stream_>>>(d_var_lay_begin_.p, d_var_lay_.p, d_mmd_.p, delta_out, (uint32_t)n_vars_);
            ++launches_;
            CUDA_CHECK(cudaGetLastError());
        }
    }

    void forward_pass(double omega) override
    {
        set_device();
        if(!backward_valid_) backward_run();     // bdd_cuda_parallel_mma.cu:210-211
        REAL* in = dbuf(dcur_);
        REAL* out = dbuf((dcur_ + 1) % 3);
        REAL* zero = dbuf((dcur_ + 2) % 3);
        mma_pass<true>(omega, in, delta_needs_norm_, out, zero);
        dcur_ = (dcur_ + 1) % 3; delta_needs_norm_ = true;
        forward_valid_ = true; backward_valid_ = false;
        if(xc_.mode != 0) launch_exchange();
    }

    // ---- multi-GPU: the exchange of the shared variables' sums after every pass, issued by the library itself ----------
    void set_exchange(int world, int rank, const void* const* peers, uint32_t* const* flags, void* out, void* const* outs,
                      const void* mc_in, void* mc_out, size_t n_exchange, int mode) override
    {
        set_device();
        if(graph_exec_) { cudaGraphExecDestroy(graph_exec_); graph_exec_ = nullptr; }
        for(StepGraph& c : step_graphs_) cudaGraphExecDestroy(c.exec);
        step_graphs_.clear();
        if(mode == 0) { xc_ = Xchg{}; return; }
        if(world < 2 || world > EXCHANGE_MAX_WORLD || rank < 0 || rank >= world || flags == nullptr || (n_exchange & 1) || n_exchange > 2 * n_vars_ || ext_delta_[0] == nullptr)
            throw api_error(BDDB200_ERR_INVALID_ARGUMENT, "set_exchange: invalid argument (the sum buffers must be set with set_delta_buffers first)");
        if((mode == 1 && (peers == nullptr || out == nullptr)) || (mode == 2 && (peers == nullptr || outs == nullptr)) || (mode == 3 && (mc_in == nullptr || mc_out == nullptr))
           || (mode == 4 && peers == nullptr) || mode < 0 || mode > 4)
            throw api_error(BDDB200_ERR_INVALID_ARGUMENT, "set_exchange: buffers missing for the requested mode");
        if(mode == 4 && (deterministic_ || n_lane_ != n_bundles_ || delta_in_override_ != nullptr))
            throw api_error(BDDB200_ERR_STATE, "set_exchange: the push exchange needs the default (atomic) sums, a collection of lane-class bundles only and no separate input buffer");
        xc_.world = world; xc_.rank = rank; xc_.peers = peers; xc_.flags = flags; xc_.out = static_cast<REAL*>(out); xc_.outs = outs;
        xc_.mc_in = static_cast<const REAL*>(mc_in); xc_.mc_out = static_cast<REAL*>(mc_out); xc_.n_exchange = n_exchange; xc_.mode = mode;
        if(d_xc_counters_.n < 8 + 2 * PUSH_SLOTS) d_xc_counters_.alloc(8 + 2 * PUSH_SLOTS);
        d_xc_counters_.zero(stream_);
        xc_phase_ = 0; push_pending_ = false;
        if(mode == 4)
        {   // this rank's own flag array (the passes poll it; peers write it), and which bundles take part in the flag barrier
            CUDA_CHECK(cudaMemcpyAsync(&xc_.my_flags, flags + rank, sizeof(uint32_t*), cudaMemcpyDeviceToHost, stream_));
            // which other ranks hold each shared variable (bddb200_create_shard / bddb200_set_push_masks); without that knowledge: all of them
            {
                const size_t n_sh = n_exchange / 2;
                std::vector<uint16_t> m(std::max<size_t>(n_sh, 1), (uint16_t)(((1u << world) - 1u) & ~(1u << rank)));
                if(h_push_mask_.size() >= n_sh) for(size_t v = 0; v < n_sh; ++v) m[v] = (uint16_t)(h_push_mask_[v] & ((1u << world) - 1u) & ~(1u << rank));
                d_push_mask_.upload(m, stream_);
            }
            // the bundles that take part in the pass-end barrier: those with shared variables -- the first n_lane_shared_ when the layout
            // was built knowing the shared prefix (bddb200_options.n_shared_vars), else all of them -- and those that clear the prefix
            push_n_shared_bundles_ = (layout_shared_vars_ == n_exchange / 2) ? n_lane_shared_ : n_lane_;
            push_debug_ = 0;
            if(const char* e = std::getenv("BDDB200_PUSH_DEBUG")) push_debug_ = (uint32_t)std::atoi(e);      // diagnostics only (kernels.cuh: SweepArgs::push_debug)
            const size_t n_with_shared = push_n_shared_bundles_;
            const size_t zpb = (n_vars_ + n_lane_ - 1) / n_lane_;
            std::vector<uint32_t> h_cnt(8 + 2 * PUSH_SLOTS, 0u);
            for(size_t g = 0; g < n_lane_; ++g)
                if(g < n_with_shared || g == 0 || g * zpb < n_exchange / 2) ++h_cnt[8 + PUSH_SLOTS + g % PUSH_SLOTS];
            for(int sl = 0; sl < PUSH_SLOTS; ++sl) h_cnt[3] += h_cnt[8 + PUSH_SLOTS + sl] != 0;
            CUDA_CHECK(cudaMemcpyAsync(d_xc_counters_.p, h_cnt.data(), h_cnt.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream_));
            CUDA_CHECK(cudaStreamSynchronize(stream_));
        }
    }
    // push exchange: the sums of the last pass are complete on this rank once every peer has finished that pass (what = 1); before the
    // peers may push into buffers this rank has cleared outside a pass they must hear about it (what = 2)
    void push_barrier(int what)
    {
        if(xc_.mode != 4) return;
        if(what == 1) { if(!push_pending_) return; push_pending_ = false; }       // nothing pushed since the last wait
        const uint32_t stale = (xc_phase_ + 2u) % 3u, send = (xc_phase_ + 1u) % 3u;
        push_barrier_kernel<<<1, 32, 0, stream_>>>(d_xc_counters_.p, xc_.flags, xc_.my_flags, xc_.world, xc_.rank, what, stale, send);
        if(what & 2) xc_phase_ = send;
        ++launches_;
        CUDA_CHECK(cudaGetLastError());
    }
    void launch_exchange()
    {
        if(xc_.n_exchange == 0 || xc_.mode == 4) return;       // push exchange: done by the pass itself
        const size_t offset = (size_t)(dbuf(dcur_) - dbuf(0));       // the sums of the pass just made, relative to the start of the symmetric block
        const size_t pairs = xc_.n_exchange / 2;
        // programmatic dependent launch, like the sweeps: the exchange sets itself up under the tail of the pass and the next pass under the exchange
        cudaLaunchConfig_t cfg{};
        cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = stream_;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = pdl_ ? 1 : 0;
        if(xc_.mode == 1)
        {
            cfg.gridDim = dim3((unsigned)std::max<size_t>(1, std::min<size_t>((pairs + 255) / 256, (size_t)n_sms_ * 4)));
            CUDA_CHECK(cudaLaunchKernelEx(&cfg, delta_exchange_kernel<REAL>, reinterpret_cast<const REAL* const*>(xc_.peers), xc_.flags, xc_.world, xc_.rank, 0u, offset,
                                          xc_.out, pairs, d_xc_counters_.p));
        }
        else if(xc_.mode == 2)
        {
            const size_t per = (pairs + xc_.world - 1) / xc_.world;
            cfg.gridDim = dim3((unsigned)std::max<size_t>(1, std::min<size_t>((per + 255) / 256, (size_t)n_sms_ * 4)));
            CUDA_CHECK(cudaLaunchKernelEx(&cfg, delta_exchange2_kernel<REAL>, reinterpret_cast<const REAL* const*>(xc_.peers), reinterpret_cast<REAL* const*>(xc_.outs), xc_.flags,
                                          xc_.world, xc_.rank, 0u, offset, pairs, d_xc_counters_.p));
        }
        else
        {
            const size_t units = (xc_.n_exchange * sizeof(REAL) + 15) / 16, per = (units + xc_.world - 1) / xc_.world;
            cfg.gridDim = dim3((unsigned)std::max<size_t>(1, std::min<size_t>((per + 255) / 256, (size_t)n_sms_ * 2)));
            CUDA_CHECK(cudaLaunchKernelEx(&cfg, delta_exchange_mc_kernel<REAL>, xc_.mc_in, (REAL*)xc_.mc_out, xc_.flags, xc_.world, xc_.rank, offset, xc_.n_exchange, d_xc_counters_.p));
        }
        ++launches_;
        CUDA_CHECK(cudaGetLastError());
    }

    void backward_pass(double omega) override
    {
        set_device();
        if(!forward_valid_) throw api_error(BDDB200_ERR_STATE, "backward_mm needs a valid forward state (call forward_mm first)");
        REAL* in = dbuf(dcur_);
        REAL* out = dbuf((dcur_ + 1) % 3);
        REAL* zero = dbuf((dcur_ + 2) % 3);
        mma_pass<false>(omega, in, delta_needs_norm_, out, zero);
        dcur_ = (dcur_ + 1) % 3; delta_needs_norm_ = true;
        forward_valid_ = false; backward_valid_ = true; lb_valid_ = false;
        if(xc_.mode != 0) launch_exchange();
    }

    void iteration(double omega) override
    {
        if(use_resident()) { set_device(); launch_resident(omega, 1); return; }
        forward_pass(omega);
        backward_pass(omega);
    }

    void iterations(double omega, size_t n) override
    {
        set_device();
        if(n == 0) return;
        if(use_resident()) { launch_resident(omega, n); return; }
        if(!backward_valid_) backward_run();
        // The delta buffers rotate with period 3 passes and an iteration is 2 passes: a graph of
        // 3 iterations (6 sweep launches) returns to the same buffer assignment.
        const size_t per_graph = 3;
        if(n >= per_graph)
        {
            if(graph_exec_ == nullptr || graph_omega_ != omega || graph_dcur_ != dcur_ || graph_cc_ != cc_ || graph_norm_ != delta_needs_norm_ || graph_phase_ != xc_phase_)
            {
                if(graph_exec_) { cudaGraphExecDestroy(graph_exec_); graph_exec_ = nullptr; }
                if(!delta_needs_norm_)
                {   // first passes read a normalised vector: run one plain iteration so that the
                    // captured graph always starts from the steady state
                    iteration(omega); --n;
                }
            }
            if(n >= per_graph && graph_exec_ == nullptr)
            {
                cudaGraph_t graph = nullptr;
                const size_t launches_before = launches_;
                CUDA_CHECK(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal));
                for(size_t i = 0; i < per_graph; ++i) iteration(omega);
                CUDA_CHECK(cudaStreamEndCapture(stream_, &graph));
                graph_launches_ = launches_ - launches_before;
                launches_ = launches_before;
                CUDA_CHECK(cudaGraphInstantiate(&graph_exec_, graph, 0));
                cudaGraphDestroy(graph);
                graph_omega_ = omega; graph_dcur_ = dcur_; graph_cc_ = cc_; graph_norm_ = delta_needs_norm_; graph_phase_ = xc_phase_;     // six passes: the phase is back where it was
            }
            while(n >= per_graph && graph_exec_ != nullptr)
            {
                CUDA_CHECK(cudaGraphLaunch(graph_exec_, stream_));
                launches_ += graph_launches_;
                n -= per_graph;
            }
            forward_valid_ = false; backward_valid_ = true; lb_valid_ = false;
        }
        for(size_t i = 0; i < n; ++i) iteration(omega);
    }

    // forward_mm(omega, delta) on a caller-owned vector: in = values to add (already
    // normalised by the caller), out = raw sums (bdd_cuda_parallel_mma.cu:207-257).
    void forward_mm(double omega, void* delta_dev) override
    {
        set_device();
        if(!backward_valid_) backward_run();
        external_pass<true>(omega, static_cast<REAL*>(delta_dev));
        forward_valid_ = true; backward_valid_ = false;
    }
    void backward_mm(double omega, void* delta_dev) override
    {
        set_device();
        if(!forward_valid_) throw api_error(BDDB200_ERR_STATE, "backward_mm needs a valid forward state (call forward_mm first)");
        external_pass<false>(omega, static_cast<REAL*>(delta_dev));
        forward_valid_ = false; backward_valid_ = true; lb_valid_ = false;
    }
    template<bool FORWARD>
    void external_pass(double omega, REAL* delta)
    {
        REAL* out = d_delta_tmp_.p;
        if(!deterministic_) d_delta_tmp_.zero(stream_);
        mma_pass<FORWARD>(omega, delta, false, out, nullptr);
        CUDA_CHECK(cudaMemcpyAsync(delta, out, sizeof(REAL) * 2 * n_vars_, cudaMemcpyDeviceToDevice, stream_));
    }

    void normalize_delta(void* delta_dev) const override
    {
        set_device();
        REAL* d = static_cast<REAL*>(delta_dev);
        normalize_kernel<REAL><<<blocks_for(2 * n_vars_), 256, 0, stream_>>>(d, d, d_nr_bdds_.p, (uint32_t)(2 * n_vars_));
        ++launches_;
        CUDA_CHECK(cudaGetLastError());
    }

    void get_delta(void* out, int out_is_host) override
    {
        set_device();
        ensure_sums();
        push_barrier(1);
        const REAL* src = dbuf(dcur_);
        if(delta_in_override_ && n_shared_vars_ > 0)
        {   // exchanged sums of the shared variables, local sums of the rest
            CUDA_CHECK(cudaMemcpyAsync(d_delta_tmp2_.p, src, sizeof(REAL) * 2 * n_vars_, cudaMemcpyDeviceToDevice, stream_));
            CUDA_CHECK(cudaMemcpyAsync(d_delta_tmp2_.p, delta_in_override_, sizeof(REAL) * 2 * n_shared_vars_, cudaMemcpyDeviceToDevice, stream_));
            src = d_delta_tmp2_.p;
        }
        if(delta_needs_norm_)
        {
            normalize_kernel<REAL><<<blocks_for(2 * n_vars_), 256, 0, stream_>>>(src, d_delta_tmp_.p, d_nr_bdds_.p, (uint32_t)(2 * n_vars_));
            ++launches_;
            src = d_delta_tmp_.p;
        }
        CUDA_CHECK(cudaMemcpyAsync(out, src, sizeof(REAL) * 2 * n_vars_, out_is_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, stream_));
        if(out_is_host) CUDA_CHECK(cudaStreamSynchronize(stream_));
    }

    void* delta_sum_buffer() override { set_device(); ensure_sums(); push_barrier(1); return dbuf(dcur_); }
    int delta_sum_index() const override { return dcur_; }
    void set_push_masks(const uint16_t* masks_host, size_t n) override { h_push_mask_.assign(masks_host, masks_host + n); }
    int push_exchange_supported() const override { return !deterministic_ && n_lane_ == n_bundles_ ? 1 : 0; }
    // Multi-GPU exchange over peer memory: the three rotating sum buffers live in caller-owned (symmetric) memory, and the
    // passes read the exchanged sums from a separate buffer (bddb200_delta_exchange writes it).
    void set_delta_buffers(void* b0, void* b1, void* b2) override
    {
        if(graph_exec_) { cudaGraphExecDestroy(graph_exec_); graph_exec_ = nullptr; }
        for(StepGraph& c : step_graphs_) cudaGraphExecDestroy(c.exec);
        step_graphs_.clear();
        ext_delta_[0] = static_cast<REAL*>(b0); ext_delta_[1] = static_cast<REAL*>(b1); ext_delta_[2] = static_cast<REAL*>(b2);
    }
    void set_delta_input(void* in, size_t n_shared_vars) override
    {
        if(n_shared_vars > n_vars_) throw api_error(BDDB200_ERR_INVALID_ARGUMENT, "more shared variables than variables");
        if(graph_exec_) { cudaGraphExecDestroy(graph_exec_); graph_exec_ = nullptr; }
        for(StepGraph& c : step_graphs_) cudaGraphExecDestroy(c.exec);
        step_graphs_.clear();
        delta_in_override_ = static_cast<REAL*>(in);
        n_shared_vars_ = in ? n_shared_vars : 0;
        if(in && d_delta_tmp2_.n == 0) d_delta_tmp2_.alloc(2 * n_vars_);
    }

    // diagnostics: one MMA pass with per-bundle clock64() stamps (kernels.cuh, TRACE_EVENTS per bundle)
    size_t trace_pass(int forward, double omega, unsigned long long* out_host, size_t max_bundles) override
    {
        set_device();
        const size_t n = std::min(max_bundles, n_lane_ > 0 ? n_lane_ : n_small_);      // bundles of the first launch
        DevBuf<unsigned long long> tr; tr.alloc((n_bundles_ + 64) * TRACE_EVENTS); tr.zero(stream_);
        if(forward == 2)
        {   // one iteration of the on-chip kernel
            if(!use_resident()) throw api_error(BDDB200_ERR_STATE, "trace_pass(2): the on-chip kernel is not in use for this solver");
            launch_resident(omega, 1, tr.p);
            CUDA_CHECK(cudaMemcpyAsync(out_host, tr.p, n * TRACE_EVENTS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream_));
            CUDA_CHECK(cudaStreamSynchronize(stream_));
            return n;
        }
        trace_ = tr.p;
        if(forward) forward_pass(omega); else backward_pass(omega);
        trace_ = nullptr;
        CUDA_CHECK(cudaMemcpyAsync(out_host, tr.p, n * TRACE_EVENTS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream_));
        CUDA_CHECK(cudaStreamSynchronize(stream_));
        return n;
    }

    // ---------------------------------------------------------------- plain runs -------
    void forward_run() override
    {
        set_device();
        if(forward_valid_) return;
        SweepArgs<REAL> a = base_args();
        launch_sweep<MODE_PLAIN, true>(a);
        forward_valid_ = true;
    }
    void backward_run() override
    {
        set_device();
        if(backward_valid_) return;
        SweepArgs<REAL> a = base_args();
        launch_sweep<MODE_PLAIN, false>(a);
        backward_valid_ = true; lb_valid_ = false;
    }
    void flush_forward() override { forward_valid_ = false; }
    void flush_backward() override { backward_valid_ = false; lb_valid_ = false; }

    double lower_bound() override
    {
        set_device();
        if(xc_.mode == 4) push_barrier(1);           // a natural end point of a sharded solve: the peers are done adding to this rank's sums
        if(xc_.mode != 0 && !lb_valid_) check_exchange_status();
        backward_run();
        if(!lb_valid_)
        {
            const double* src = d_lb_partial_.p + LB_BLOCKS + 1;      // accumulated by the backward sweep itself
            if(lb_from_resident_ && !deterministic_)
            {   // one partial sum per bundle from the on-chip kernel's last backward pass, added in bundle order
                CUDA_CHECK(cudaMemcpyAsync(h_lb_part_, d_lb_part_.p, n_lane_ * sizeof(double), cudaMemcpyDeviceToHost, stream_));
                CUDA_CHECK(cudaStreamSynchronize(stream_));
                lb_ = 0.0;
                for(size_t i = 0; i < n_lane_; ++i) lb_ += h_lb_part_[i];
                lb_valid_ = true;
                return lb_;
            }
            if(deterministic_)
            {   // fixed-shape two-stage tree over the per-BDD values (bit-reproducible)
                lb_partial_kernel<REAL><<<LB_BLOCKS, 256, 0, stream_>>>(d_bdd_lb_.p, d_lb_partial_.p, (uint32_t)n_bdds_);
                lb_final_kernel<<<1, 256, 0, stream_>>>(d_lb_partial_.p, d_lb_partial_.p + LB_BLOCKS, LB_BLOCKS);
                launches_ += 2;
                src = d_lb_partial_.p + LB_BLOCKS;
            }
            const int n_parts = deterministic_ ? 1 : LB_SLOTS;          // the sweep keeps LB_SLOTS partial sums (one per CTA index mod LB_SLOTS)
            CUDA_CHECK(cudaMemcpyAsync(h_lb_, src, n_parts * sizeof(double), cudaMemcpyDeviceToHost, stream_));
            CUDA_CHECK(cudaStreamSynchronize(stream_));
            lb_ = 0.0;
            for(int i = 0; i < n_parts; ++i) lb_ += h_lb_[i];
            lb_valid_ = true;
        }
        return lb_;
    }

    void lower_bound_per_bdd(void* out_dev) override
    {
        set_device();
        backward_run();
        CUDA_CHECK(cudaMemcpyAsync(out_dev, d_bdd_lb_.p, sizeof(REAL) * n_bdds_, cudaMemcpyDeviceToDevice, stream_));
    }

    // ---------------------------------------------------------------- costs -------------
    // Host cost vectors go through a pinned staging buffer (converted to REAL on the way), one
    // H2D copy and one kernel, all asynchronous; the caller's arrays are free on return.
    template<typename SRC>
    void update_costs_host_impl(const SRC* lo, size_t n_lo, const SRC* hi, size_t n_hi)
    {
        set_device();
        if(n_lo > n_vars_ || n_hi > n_vars_) throw api_error(BDDB200_ERR_INVALID_ARGUMENT, "more costs than variables");
        if(n_lo + n_hi == 0) return;
        if(h_stage_ == nullptr)
        {
            CUDA_CHECK(cudaMallocHost(&h_stage_, sizeof(REAL) * 2 * std::max<size_t>(n_vars_, 1)));
            d_stage_.alloc(2 * n_vars_);
            CUDA_CHECK(cudaEventCreateWithFlags(&stage_free_, cudaEventDisableTiming));
        }
        else CUDA_CHECK(cudaEventSynchronize(stage_free_));     // previous upload has left the staging buffer
        if(std::is_same<SRC, REAL>::value)
        {
            if(n_lo) std::memcpy(h_stage_, lo, n_lo * sizeof(REAL));
            if(n_hi) std::memcpy(h_stage_ + n_lo, hi, n_hi * sizeof(REAL));
        }
        else
        {
            for(size_t i = 0; i < n_lo; ++i) h_stage_[i] = (REAL)lo[i];          // device_vector<REAL>(cost_begin, cost_end), bdd_cuda_base.cu:485
            for(size_t i = 0; i < n_hi; ++i) h_stage_[n_lo + i] = (REAL)hi[i];
        }
        CUDA_CHECK(cudaMemcpyAsync(d_stage_.p, h_stage_, sizeof(REAL) * (n_lo + n_hi), cudaMemcpyHostToDevice, stream_));
        CUDA_CHECK(cudaEventRecord(stage_free_, stream_));
        update_costs_lohi_kernel<REAL><<<blocks_for(n_lay_), 256, 0, stream_>>>(d_lay_vn_.p, d_lohi_[cc_].p, d_stage_.p, (uint32_t)n_lo, (uint32_t)n_hi, (uint32_t)n_lay_);
        ++launches_;
        CUDA_CHECK(cudaGetLastError());
        flush_forward(); flush_backward();
    }
    void update_costs_host(const double* lo, size_t n_lo, const double* hi, size_t n_hi) override { update_costs_host_impl<double>(lo, n_lo, hi, n_hi); }
    void update_costs_host_real(const void* lo, size_t n_lo, const void* hi, size_t n_hi) override
    { update_costs_host_impl<REAL>(static_cast<const REAL*>(lo), n_lo, static_cast<const REAL*>(hi), n_hi); }
    void update_costs_dev(const void* lo, size_t n_lo, const void* hi, size_t n_hi) override
    {
        set_device();
        if((n_lo != 0 && n_lo != n_vars_) || (n_hi != 0 && n_hi != n_vars_))
            throw api_error(BDDB200_ERR_INVALID_ARGUMENT, "device cost vectors must have 0 or nr_variables entries");  // bdd_cuda_base.cu:537-538
        if(n_lo) { update_costs_kernel<REAL><<<blocks_for(n_lay_), 256, 0, stream_>>>(d_lay_vn_.p, d_lohi_[cc_].p, static_cast<const REAL*>(lo), (uint32_t)n_lo, (uint32_t)n_lay_); ++launches_; }
        if(n_hi) { update_costs_kernel<REAL><<<blocks_for(n_lay_), 256, 0, stream_>>>(d_lay_vn_.p, d_lohi_[cc_].p + 1, static_cast<const REAL*>(hi), (uint32_t)n_hi, (uint32_t)n_lay_); ++launches_; }
        CUDA_CHECK(cudaGetLastError());
        flush_forward(); flush_backward();
    }
    // One solver step of a host-driven loop in ONE call: update_costs(lo, hi) from host vectors, iteration(omega), lower_bound()
    // (what the perturbation rounds of bdd_solver.cpp:318-380 and run_solver, run_solver_util.h:37-49, do per step).  Vectors of REAL
    // in pinned memory are read by the copy engine straight from the caller's buffers (anything else goes through the pinned staging
    // buffer); upload, cost update, the plain backward sweep the changed costs call for, both MMA sweeps and the read-back of the
    // bound are ONE CUDA graph launch (one graph per host buffer and rotation state, cached).  The call returns after the bound has
    // come back, so the caller's buffers are free again.
    double step_host(const void* lo, size_t n_lo, const void* hi, size_t n_hi, int src_is_real, double omega) override
    {
        set_device();
        if(n_lo > n_vars_ || n_hi > n_vars_) throw api_error(BDDB200_ERR_INVALID_ARGUMENT, "more costs than variables");
        if(use_resident() || deterministic_)
        {   // these forms have their own launch structure: plain sequence
            if(src_is_real) update_costs_host_real(lo, n_lo, hi, n_hi); else update_costs_host(static_cast<const double*>(lo), n_lo, static_cast<const double*>(hi), n_hi);
            iteration(omega);
            return lower_bound();
        }
        // where the upload reads from: the caller's own memory when it is pinned REAL data, else the pinned staging buffer
        const void* src_lo = lo; const void* src_hi = hi;
        if(n_lo + n_hi > 0)
        {
            if(h_stage_ == nullptr)
            {
                CUDA_CHECK(cudaMallocHost(&h_stage_, sizeof(REAL) * 2 * std::max<size_t>(n_vars_, 1)));
                d_stage_.alloc(2 * n_vars_);
                CUDA_CHECK(cudaEventCreateWithFlags(&stage_free_, cudaEventDisableTiming));
            }
            auto pinned = [](const void* p) {
                if(p == nullptr) return true;
                cudaPointerAttributes at{};
                if(cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
                return at.type == cudaMemoryTypeHost;
            };
            if(!(src_is_real && pinned(n_lo ? lo : nullptr) && pinned(n_hi ? hi : nullptr)))
            {   // the previous step has been waited for, so the staging buffer is free
                if(src_is_real)
                {
                    if(n_lo) std::memcpy(h_stage_, lo, n_lo * sizeof(REAL));
                    if(n_hi) std::memcpy(h_stage_ + n_lo, hi, n_hi * sizeof(REAL));
                }
                else
                {
                    const double* l = static_cast<const double*>(lo); const double* h = static_cast<const double*>(hi);
                    for(size_t i = 0; i < n_lo; ++i) h_stage_[i] = (REAL)l[i];
                    for(size_t i = 0; i < n_hi; ++i) h_stage_[n_lo + i] = (REAL)h[i];
                }
                src_lo = h_stage_; src_hi = h_stage_ + n_lo;
            }
        }
        ensure_sums();
        StepKey key{dcur_, cc_, delta_needs_norm_, omega, n_lo, n_hi, n_lo ? src_lo : nullptr, n_hi ? src_hi : nullptr, xc_phase_};
        StepGraph* g = nullptr;
        for(StepGraph& c : step_graphs_) if(c.key == key) g = &c;
        if(g == nullptr)
        {
            if(step_graphs_.size() >= 24) { for(StepGraph& c : step_graphs_) cudaGraphExecDestroy(c.exec); step_graphs_.clear(); }
            cudaGraph_t graph = nullptr;
            const size_t launches_before = launches_;
            const int dcur0 = dcur_, cc0 = cc_; const bool norm0 = delta_needs_norm_; const uint32_t phase0 = xc_phase_;
            lb_sum_clean_ = false;           // the captured backward sweep clears the partial sums itself, whatever ran before a replay
            CUDA_CHECK(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal));
            if(n_lo + n_hi > 0)
            {   // one copy when the two vectors are adjacent in host memory
                if(n_lo && n_hi && static_cast<const REAL*>(src_lo) + n_lo == static_cast<const REAL*>(src_hi))
                    CUDA_CHECK(cudaMemcpyAsync(d_stage_.p, src_lo, sizeof(REAL) * (n_lo + n_hi), cudaMemcpyHostToDevice, stream_));
                else
                {
                    if(n_lo) CUDA_CHECK(cudaMemcpyAsync(d_stage_.p, src_lo, sizeof(REAL) * n_lo, cudaMemcpyHostToDevice, stream_));
                    if(n_hi) CUDA_CHECK(cudaMemcpyAsync(d_stage_.p + n_lo, src_hi, sizeof(REAL) * n_hi, cudaMemcpyHostToDevice, stream_));
                }
            }
            step_body(n_lo, n_hi, omega);
            CUDA_CHECK(cudaStreamEndCapture(stream_, &graph));
            StepGraph ng; ng.key = key; ng.launches = launches_ - launches_before;
            launches_ = launches_before;
            CUDA_CHECK(cudaGraphInstantiate(&ng.exec, graph, 0));
            cudaGraphDestroy(graph);
            step_graphs_.push_back(ng);
            g = &step_graphs_.back();
            dcur_ = dcur0; cc_ = cc0; delta_needs_norm_ = norm0; xc_phase_ = phase0;      // the capture only recorded the step
        }
        CUDA_CHECK(cudaGraphLaunch(g->exec, stream_));
        launches_ += g->launches;
        dcur_ = (dcur_ + 2) % 3; delta_needs_norm_ = true;          // two passes; the cost buffers swap twice
        if(xc_.mode == 4) xc_phase_ = (xc_phase_ + 2u) % 3u;
        forward_valid_ = false; backward_valid_ = true; lb_sum_clean_ = false; lb_from_resident_ = false;
        CUDA_CHECK(cudaStreamSynchronize(stream_));
        lb_ = 0.0;
        for(int i = 0; i < LB_SLOTS; ++i) lb_ += h_lb_[i];
        lb_valid_ = true;
        return lb_;
    }
    // what step_host captures (host state changes as in the separate calls)
    void step_body(size_t n_lo, size_t n_hi, double omega)
    {
        if(n_lo + n_hi > 0)
        {
            update_costs_lohi_kernel<REAL><<<blocks_for(n_lay_), 256, 0, stream_>>>(d_lay_vn_.p, d_lohi_[cc_].p, d_stage_.p, (uint32_t)n_lo, (uint32_t)n_hi, (uint32_t)n_lay_);
            ++launches_;
            flush_forward(); flush_backward();
        }
        forward_pass(omega);         // runs the plain backward sweep first when the costs changed
        backward_pass(omega);
        CUDA_CHECK(cudaMemcpyAsync(h_lb_, d_lb_partial_.p + LB_BLOCKS + 1, LB_SLOTS * sizeof(double), cudaMemcpyDeviceToHost, stream_));
    }

    void set_cost(double c, size_t var) override
    {
        set_device();
        if(var >= n_vars_ || h_nr_bdds_per_var_[var] <= 0) throw api_error(BDDB200_ERR_INVALID_ARGUMENT, "set_cost: variable not covered by any BDD");
        const REAL add = (REAL)c / h_nr_bdds_per_var_[var];
        set_cost_kernel<REAL><<<blocks_for(n_lay_), 256, 0, stream_>>>(d_lay_vn_.p, d_lohi_[cc_].p + 1, (int)var, add, (uint32_t)n_lay_);
        ++launches_;
        CUDA_CHECK(cudaGetLastError());
        flush_forward(); flush_backward();
    }
    void distribute_delta() override
    {
        set_device();
        distribute_kernel<REAL><<<blocks_for(n_lay_), 256, 0, stream_>>>(d_lay_vn_.p, d_lohi_[cc_].p, d_mmd_.p, (uint32_t)n_lay_);
        ++launches_;
        CUDA_CHECK(cudaGetLastError());
        push_barrier(1);        // no peer is still adding to these buffers ...
        for(int i = 0; i < 3; ++i) CUDA_CHECK(cudaMemsetAsync(dbuf(i), 0, sizeof(REAL) * 2 * n_vars_, stream_));
        if(delta_in_override_) CUDA_CHECK(cudaMemsetAsync(delta_in_override_, 0, sizeof(REAL) * 2 * n_shared_vars_, stream_));
        push_barrier(2);        // ... and none starts again before they are clear
        delta_needs_norm_ = false; exch_in_contrib_ = false;
        flush_forward(); flush_backward();
    }
    void get_solver_costs(void* lo, void* hi, void* mmd) const override
    {
        set_device();
        const unsigned nb = blocks_for(n_ext_);
        if(lo) { gather_ext_kernel<REAL><<<nb, 256, 0, stream_>>>(d_ext2lay_.p, d_ext_var_.p, d_lohi_[cc_].p, 2u, static_cast<REAL*>(lo), (REAL)0, (uint32_t)n_ext_); ++launches_; }
        if(hi) { gather_ext_kernel<REAL><<<nb, 256, 0, stream_>>>(d_ext2lay_.p, d_ext_var_.p, d_lohi_[cc_].p + 1, 2u, static_cast<REAL*>(hi), (REAL)0, (uint32_t)n_ext_); ++launches_; }
        if(mmd) { gather_ext_kernel<REAL><<<nb, 256, 0, stream_>>>(d_ext2lay_.p, d_ext_var_.p, d_mmd_.p, 1u, static_cast<REAL*>(mmd), (REAL)0, (uint32_t)n_ext_); ++launches_; }
        CUDA_CHECK(cudaGetLastError());
    }
    void set_solver_costs(const void* lo, const void* hi, const void* mmd) override
    {
        set_device();
        ensure_sums();      // the pending sums are separate state (delta_lo_hi_) in the reference
        const unsigned nb = blocks_for(n_ext_);
        scatter_ext_kernel<REAL><<<nb, 256, 0, stream_>>>(d_ext2lay_.p, d_ext_var_.p, static_cast<const REAL*>(lo), d_lohi_[cc_].p, 2u, (uint32_t)n_ext_);
        scatter_ext_kernel<REAL><<<nb, 256, 0, stream_>>>(d_ext2lay_.p, d_ext_var_.p, static_cast<const REAL*>(hi), d_lohi_[cc_].p + 1, 2u, (uint32_t)n_ext_);
        scatter_ext_kernel<REAL><<<nb, 256, 0, stream_>>>(d_ext2lay_.p, d_ext_var_.p, static_cast<const REAL*>(mmd), d_mmd_.p, 1u, (uint32_t)n_ext_);
        launches_ += 3;
        CUDA_CHECK(cudaGetLastError());
        flush_forward(); flush_backward();
    }
    void primal_objective_host(double* out) override
    {
        set_device();
        DevBuf<double> d; d.alloc(n_vars_);
        primal_objective_kernel<REAL><<<blocks_for(n_vars_), 256, 0, stream_>>>(d_var_lay_begin_.p, d_var_lay_.p, d_lohi_[cc_].p, d.p, (uint32_t)n_vars_);
        ++launches_;
        CUDA_CHECK(cudaMemcpyAsync(out, d.p, sizeof(double) * n_vars_, cudaMemcpyDeviceToHost, stream_));
        CUDA_CHECK(cudaStreamSynchronize(stream_));
    }

    // ---------------------------------------------------------------- min-marginals ----
    void min_marginals(int sorted, int32_t* primal_dev, void* lo_dev, void* hi_dev) override
    {
        set_device();
        min_marginals_internal();                        // forward_run, backward_run(true): bdd_cuda_base.cu:721-728
        const unsigned nb = blocks_for(n_ext_);
        const REAL INF = std::numeric_limits<REAL>::infinity();
        DevBuf<REAL> tmp;
        if(sorted) tmp.alloc(n_ext_);
        auto emit = [&](const REAL* src, void* dst) {
            if(dst == nullptr) return;
            REAL* stage = sorted ? tmp.p : static_cast<REAL*>(dst);
            gather_ext_kernel<REAL><<<nb, 256, 0, stream_>>>(d_ext2lay_.p, d_ext_var_.p, src, 1u, stage, INF, (uint32_t)n_ext_);
            ++launches_;
            if(sorted) { permute_kernel<REAL><<<nb, 256, 0, stream_>>>(d_sorted_ext_.p, stage, static_cast<REAL*>(dst), (uint32_t)n_ext_); ++launches_; }
        };
        emit(d_mm_lo_.p, lo_dev);
        emit(d_mm_hi_.p, hi_dev);
        if(primal_dev)
        {
            if(sorted) { permute_kernel<int32_t><<<nb, 256, 0, stream_>>>(d_sorted_ext_.p, d_ext_var_.p, primal_dev, (uint32_t)n_ext_); ++launches_; }
            else CUDA_CHECK(cudaMemcpyAsync(primal_dev, d_ext_var_.p, sizeof(int32_t) * n_ext_, cudaMemcpyDeviceToDevice, stream_));
        }
        CUDA_CHECK(cudaGetLastError());
        if(sorted) CUDA_CHECK(cudaStreamSynchronize(stream_));   // tmp is freed on return
    }

    // the same into host memory, as doubles (two_dim_variable_array<std::array<double,2>> min_marginals(), bdd_cuda_base.cu:753-786)
    void min_marginals_host(int sorted, int32_t* primal_host, double* lo_host, double* hi_host) override
    {
        set_device();
        DevBuf<int32_t> idx; DevBuf<REAL> lo, hi;
        idx.alloc(n_ext_); lo.alloc(n_ext_); hi.alloc(n_ext_);
        min_marginals(sorted, idx.p, lo.p, hi.p);
        std::vector<REAL> h(n_ext_);
        if(primal_host) CUDA_CHECK(cudaMemcpyAsync(primal_host, idx.p, sizeof(int32_t) * n_ext_, cudaMemcpyDeviceToHost, stream_));
        for(int k = 0; k < 2; ++k)
        {
            double* dst = k == 0 ? lo_host : hi_host;
            if(dst == nullptr) continue;
            CUDA_CHECK(cudaMemcpyAsync(h.data(), (k == 0 ? lo : hi).p, sizeof(REAL) * n_ext_, cudaMemcpyDeviceToHost, stream_));
            CUDA_CHECK(cudaStreamSynchronize(stream_));
            for(size_t i = 0; i < n_ext_; ++i) dst[i] = (double)h[i];
        }
        CUDA_CHECK(cudaStreamSynchronize(stream_));
    }

    // compute the min-marginals of every layer entry into d_mm_lo_ / d_mm_hi_ (bdd_cuda_base.cu:716-736)
    void min_marginals_internal()
    {
        forward_run();
        SweepArgs<REAL> a = base_args();
        launch_sweep<MODE_MM, false>(a);
        backward_valid_ = true; lb_valid_ = false;
    }

    // perturb_primal_costs, incremental_mm_agreement_rounding_cuda.cu:264-335.  Returns 1 (and fills sol_host) when every
    // variable's min-marginals agree on a value, else applies the perturbation with update_costs and returns 0.
    int rounding_perturb(double delta, int round_index, unsigned long long counts_out[4], char* types_dev, char* sol_host) override
    {
        set_device();
        distribute_delta();
        min_marginals_internal();
        if(d_round_types_.n == 0)
        {
            d_round_types_.alloc(n_vars_); d_round_d0_.alloc(n_vars_); d_round_d1_.alloc(n_vars_); d_round_counts_.alloc(4);
            CUDA_CHECK(cudaMallocHost(&h_round_counts_, 4 * sizeof(unsigned long long)));
        }
        d_round_counts_.zero(stream_);
        rounding_kernel<REAL><<<blocks_for(n_vars_), 256, 0, stream_>>>(d_var_lay_begin_.p, d_var_lay_.p, d_mm_lo_.p, d_mm_hi_.p, delta, (uint32_t)round_index,
                                                                       d_round_d0_.p, d_round_d1_.p, d_round_types_.p, d_round_counts_.p, (uint32_t)n_vars_);
        ++launches_;
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaMemcpyAsync(h_round_counts_, d_round_counts_.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream_));
        if(types_dev) CUDA_CHECK(cudaMemcpyAsync(types_dev, d_round_types_.p, n_vars_, cudaMemcpyDeviceToDevice, stream_));
        CUDA_CHECK(cudaStreamSynchronize(stream_));
        for(int i = 0; i < 4; ++i) counts_out[i] = h_round_counts_[i];
        if(h_round_counts_[MM_ONE] + h_round_counts_[MM_ZERO] == n_vars_)
        {   // reconstruct the solution from the min-marginals (:300-309): type one -> 1, type zero -> 0
            if(sol_host)
            {
                CUDA_CHECK(cudaMemcpyAsync(sol_host, d_round_types_.p, n_vars_, cudaMemcpyDeviceToHost, stream_));
                CUDA_CHECK(cudaStreamSynchronize(stream_));
            }
            return 1;
        }
        update_costs_dev(d_round_d0_.p, n_vars_, d_round_d1_.p, n_vars_);
        return 0;
    }

    // ---------------------------------------------------------------- L-BFGS surface ---
    void bdds_solution(char* sol_dev) override
    {
        set_device();
        forward_run();
        backward_valid_ = false;       // backward_run(true) always recomputes, bdd_cuda_base.cu:1171
        backward_run();
        bdds_solution_kernel<REAL><<<blocks_for(n_bdds_, 128), 128, 0, stream_>>>(d_bundles_.p, d_hops_.p, d_topo_.p, d_bdd_bundle_.p, d_bdd_ext_begin_.p,
            d_cfr_.p, d_cft_.p, d_lohi_[cc_].p, sol_dev, (uint32_t)n_bdds_);
        ++launches_;
        CUDA_CHECK(cudaGetLastError());
    }
    void net_solver_costs(void* out_dev) const override
    {
        set_device();
        net_costs_kernel<REAL><<<blocks_for(n_ext_), 256, 0, stream_>>>(d_ext2lay_.p, d_ext_var_.p, d_lohi_[cc_].p, d_mmd_.p, static_cast<REAL*>(out_dev), (uint32_t)n_ext_);
        ++launches_;
        CUDA_CHECK(cudaGetLastError());
    }
    void make_dual_feasible(void* inout_dev) const override
    {
        set_device();
        REAL* d = static_cast<REAL*>(inout_dev);
        dual_feasible_kernel<REAL><<<blocks_for(n_vars_), 256, 0, stream_>>>(d_var_lay_begin_.p, d_sorted_ext_.p, d_nr_bdds_.p, d, (uint32_t)n_vars_);
        zero_terminals_kernel<REAL><<<blocks_for(n_ext_), 256, 0, stream_>>>(d_ext_var_.p, d, (uint32_t)n_ext_);
        launches_ += 2;
        CUDA_CHECK(cudaGetLastError());
    }
    void gradient_step(const void* dir_dev, double step) override
    {
        set_device();
        gradient_step_kernel<REAL><<<blocks_for(n_ext_), 256, 0, stream_>>>(d_ext2lay_.p, d_ext_var_.p, d_lohi_[cc_].p, static_cast<const REAL*>(dir_dev), (REAL)step, (uint32_t)n_ext_);
        ++launches_;
        CUDA_CHECK(cudaGetLastError());
        flush_forward(); flush_backward();
    }

    void synchronize() override { set_device(); CUDA_CHECK(cudaStreamSynchronize(stream_)); check_exchange_status(); }
    // a peer did not arrive within the exchange kernels' time limit (kernels.cuh, EXCHANGE_TIMEOUT_NS)
    void check_exchange_status()
    {
        if(xc_.mode == 0 || d_xc_counters_.n == 0) return;
        uint32_t st = 0;
        CUDA_CHECK(cudaMemcpyAsync(&st, d_xc_counters_.p + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream_));
        CUDA_CHECK(cudaStreamSynchronize(stream_));
        if(st != 0) throw api_error(BDDB200_ERR_EXCHANGE, "multi-GPU exchange: a peer did not reach the exchange within the time limit (the sums of this pass are incomplete)");
    }
    void* stream_handle() override { return (void*)stream_; }
    size_t kernel_launches() const override { return launches_; }

private:
    void set_device() const { CUDA_CHECK(cudaSetDevice(device)); }
    REAL* dbuf(int i) const { return ext_delta_[i] ? ext_delta_[i] : d_delta_[i].p; }
    REAL* ext_delta_[3] = {nullptr, nullptr, nullptr};
    struct Xchg {
        int world = 0, rank = 0, mode = 0;         // mode: 0 off, 1 one-shot, 2 two-shot, 3 in-switch (multicast), 4 push (multimem.red inside the pass)
        uint32_t* my_flags = nullptr;
        const void* const* peers = nullptr; uint32_t* const* flags = nullptr; void* const* outs = nullptr;
        REAL* out = nullptr; const REAL* mc_in = nullptr; REAL* mc_out = nullptr; size_t n_exchange = 0;
    } xc_;
    DevBuf<uint32_t> d_xc_counters_;         // {exchanges completed, CTAs finished, error, ...}: the device-side epoch of the exchange kernels; push exchange: CTA counts
    uint32_t xc_phase_ = 0;                  // push exchange: number of push barriers issued so far, mod 3
    std::vector<uint16_t> h_push_mask_;      // push exchange: per shared variable the ranks whose shards contain it (bit r = rank r), from the shard plan
    DevBuf<uint16_t> d_push_mask_;           // ... without this rank's own bit
    uint32_t push_debug_ = 0;
    bool push_pending_ = false;              // a push pass ran since this rank last waited for the peers' flags
    size_t n_lane_shared_ = 0, layout_shared_vars_ = 0, push_n_shared_bundles_ = 0;      // shard mode: lane-class bundles [0, n_lane_shared_) contain a variable < layout_shared_vars_
    REAL* delta_in_override_ = nullptr;      // exchanged sums of the variables [0, n_shared_vars_)
    size_t n_shared_vars_ = 0;
    DevBuf<REAL> d_delta_tmp2_;

    static constexpr unsigned LB_BLOCKS = 128;
    cudaStream_t stream_ = nullptr;
    bool own_stream_ = false;
    bool deterministic_ = false;
    int n_sms_ = 0, max_optin_ = 0;
    size_t n_lane_ = 0;
    uint32_t lane_max_J_ = 0, lane_max_hops_ = 0, lane_chunk_hops_ = 1, lane_stages_ = 2, lane_stage_bytes_ = 0, lane_warp_smem_ = 0;
    unsigned lane_wpc_ = 1, lane_grid_ = 1;
    bool lane_dense_ = false;          // MMA passes use the register-capped kernel (24 warps per SM)
    std::vector<LaneDesc> lane_cls_first_;
    std::vector<uint32_t> lane_cls_begin_;
    unsigned warps_per_cta_ = 4, grid_small_ = 1, forced_wpc_ = 0, n_stages_ = 3, n_stages_large_ = 2;
    uint32_t stage_small_ = 0, stage_large_ = 0, warp_smem_small_ = 0, warp_smem_large_ = 0;
    size_t n_vars_ = 0, n_bdds_ = 0, n_instr_ = 0, n_ext_ = 0, n_slots_ = 0, n_lay_ = 0, max_hops_ = 0, n_bundles_ = 0, n_small_ = 0;
    uint32_t tile_small_ = 32, tile_large_ = 0;
    std::vector<int32_t> h_nr_bdds_per_var_, h_ext_var_, h_ext_bdd_;

    DevBuf<BundleDesc> d_bundles_;
    DevBuf<ChunkRec> d_chunks_;
    DevBuf<uint32_t> d_desc_fwd_, d_desc_bwd_;
    DevBuf<LaneDesc> d_desc_lane_;
    DevBuf<REAL> d_inv_tab_;
    uint32_t inv_count_ = 1;
    bool pdl_ = true;
    DevBuf<HopRec> d_hops_;
    DevBuf<uint32_t> d_topo_, d_bdd_bundle_, d_ext2lay_, d_bdd_ext_begin_, d_var_lay_begin_, d_var_lay_, d_sorted_ext_;
    DevBuf<int2> d_lay_vn_;
    DevBuf<int32_t> d_bundle_bdd_, d_ext_var_, d_ext_bdd_, d_nr_bdds_;
    DevBuf<REAL> d_cfr_, d_cft_, d_lohi_[2], d_mmd_, d_mm_lo_, d_mm_hi_, d_delta_[3], d_delta_tmp_, d_bdd_lb_;
    DevBuf<double> d_lb_partial_;
    DevBuf<unsigned char> d_contrib_, d_sums_;      // exchange records of the on-chip kernel (resident.cuh)
    DevBuf<double> d_lb_part_;
    double* h_lb_part_ = nullptr;
    uint32_t res_vars_per_bundle_ = 0, res_own_cap_ = 0, res_own_off_ = 0;
    std::vector<uint32_t> L_var_lay_begin_;         // host copy, alive during planning only
    uint32_t res_pass_ = 0;              // number of the last pass of the on-chip kernel
    bool exch_in_contrib_ = false;       // the pending sums exist as contribution records, the rotating sum buffer is stale
    bool lb_from_resident_ = false;      // the last backward sweep was the on-chip kernel's: the bound is in d_lb_part_
    bool res_ok_ = false;
    unsigned res_grid_ = 1, res_wpc_ = 1;
    uint32_t res_warp_smem_ = 0;
    DevBuf<char> d_round_types_;
    DevBuf<REAL> d_round_d0_, d_round_d1_;
    DevBuf<unsigned long long> d_round_counts_;
    unsigned long long* h_round_counts_ = nullptr;
    double* h_lb_ = nullptr;
    REAL* h_stage_ = nullptr;            // pinned staging of host cost vectors
    DevBuf<REAL> d_stage_;
    cudaEvent_t stage_free_ = nullptr;

    unsigned long long* trace_ = nullptr;
    int cc_ = 0;                 // which of the two lo/hi cost buffers is current
    int dcur_ = 0;               // which delta buffer holds the current (pending) sums
    bool delta_needs_norm_ = false;
    bool forward_valid_ = false, backward_valid_ = false, lb_valid_ = false;
    bool lb_sum_clean_ = false;  // lb_sum holds zeros (the last sweep was a forward one, or it was just cleared)
    double lb_ = 0.0;
    mutable size_t launches_ = 0;

    struct StepKey {
        int dcur, cc; bool norm; double omega; size_t n_lo, n_hi; const void* src_lo; const void* src_hi; uint32_t phase;
        bool operator==(const StepKey& o) const
        { return dcur == o.dcur && cc == o.cc && norm == o.norm && omega == o.omega && n_lo == o.n_lo && n_hi == o.n_hi && src_lo == o.src_lo && src_hi == o.src_hi && phase == o.phase; }
    };
    struct StepGraph { StepKey key; cudaGraphExec_t exec = nullptr; size_t launches = 0; };
    std::vector<StepGraph> step_graphs_;      // step_host: one graph per rotation state of the sum buffers
    cudaGraphExec_t graph_exec_ = nullptr;
    double graph_omega_ = 0.0;
    int graph_dcur_ = 0, graph_cc_ = 0;
    uint32_t graph_phase_ = 0;
    bool graph_norm_ = false;
    size_t graph_launches_ = 0;
};

template<typename F>
int guarded(F&& f)
{
    try { f(); return BDDB200_OK; }
    catch(const layout_error& e) { g_last_error = e.what(); return e.code; }
    catch(const api_error& e) { g_last_error = e.what(); return e.code; }
    catch(const cuda_error& e) { g_last_error = e.what(); return BDDB200_ERR_CUDA; }
    catch(const std::bad_alloc&) { g_last_error = "out of host memory"; return BDDB200_ERR_INVALID_ARGUMENT; }
    catch(const std::exception& e) { g_last_error = e.what(); return BDDB200_ERR_INVALID_ARGUMENT; }
}

#define REQUIRE_SOLVER(s) if((s) == nullptr) { g_last_error = "null solver handle"; return BDDB200_ERR_INVALID_ARGUMENT; }

} // namespace

// the host-only translation units of the library (host/collection_abi.cpp) report their failures through the same string
void bddb200::detail::set_last_error(const std::string& message) { g_last_error = message; }

namespace {
// run_solver, include/run_solver_util.h:10-77, on either the plain solver or its L-BFGS wrapper
template<typename ITERATE>
void run_solver_native(bddb200_solver* s, ITERATE&& iterate, size_t max_iter, double tolerance, double improvement_slope, double time_limit_s)
{
    const auto start = std::chrono::steady_clock::now();
    const double lb_initial = s->lower_bound();
    double lb_first_iter = std::numeric_limits<double>::max(), lb_prev = lb_initial, lb_post = lb_initial;
    for(size_t it = 0; it < max_iter; ++it)
    {
        iterate();
        lb_prev = lb_post;
        lb_post = s->lower_bound();
        if(it == 0) lb_first_iter = lb_post;
        const double spent = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
        if(spent > time_limit_s) break;
        if(std::abs(lb_prev - lb_post) < std::abs(tolerance * lb_prev)) break;
        if(std::abs(lb_prev - lb_post) < improvement_slope * std::abs(lb_initial - lb_first_iter)) break;
        if(lb_post == std::numeric_limits<double>::infinity()) break;
    }
}
}

// ======================================================================== C ABI ========
extern "C" {

void bddb200_default_options(bddb200_options* o)
{
    if(!o) return;
    std::memset(o, 0, sizeof(*o));
}
const char* bddb200_last_error(void) { return g_last_error.c_str(); }
const char* bddb200_version(void) { return "bdd_b200 0.1 (sm_100a)"; }

int bddb200_create(const bddb200_instruction* instrs, size_t n_instr, const size_t* delims, size_t n_bdds,
                   const double* costs_hi, size_t n_costs, int precision, const bddb200_options* opts, bddb200_solver** out)
{
    if(out == nullptr || instrs == nullptr || delims == nullptr) { g_last_error = "null argument"; return BDDB200_ERR_INVALID_ARGUMENT; }
    *out = nullptr;
    bddb200_options o; bddb200_default_options(&o);
    if(opts) o = *opts;
    return guarded([&] {
        if(precision == BDDB200_DOUBLE) *out = new SolverImpl<double>(instrs, n_instr, delims, n_bdds, costs_hi, n_costs, o);
        else if(precision == BDDB200_FLOAT) *out = new SolverImpl<float>(instrs, n_instr, delims, n_bdds, costs_hi, n_costs, o);
        else throw api_error(BDDB200_ERR_INVALID_ARGUMENT, "precision must be BDDB200_FLOAT or BDDB200_DOUBLE");
    });
}
// ---- constraint-sharded construction: one rank's solver of a `world`-way split, planned and built in the library ----
int bddb200_plan_shard(const bddb200_instruction* instrs, size_t n_instr, const size_t* delims, size_t n_bdds, size_t nr_variables_min,
                       int world, int rank, bddb200_shard_info* info, int32_t* new_of_old_out, int32_t* counts_new_out, uint16_t* share_mask_out)
{
    (void)n_instr;
    if(instrs == nullptr || delims == nullptr || info == nullptr || n_bdds == 0) { g_last_error = "null argument"; return BDDB200_ERR_INVALID_ARGUMENT; }
    return guarded([&] {
        ShardPlan p;
        try { p = plan_shard(instrs, delims, n_bdds, nr_variables_min, world, rank); }
        catch(const std::invalid_argument& e) { throw api_error(BDDB200_ERR_INVALID_ARGUMENT, e.what()); }
        info->nr_variables = p.n_vars; info->n_shared = p.n_shared; info->shared_entries = p.shared_entries; info->first_bdd = p.first_bdd; info->n_bdds = p.n_bdds;
        if(new_of_old_out) std::memcpy(new_of_old_out, p.new_of_old.data(), p.n_vars * sizeof(int32_t));
        if(counts_new_out) std::memcpy(counts_new_out, p.counts_new.data(), p.n_vars * sizeof(int32_t));
        if(share_mask_out) std::memcpy(share_mask_out, p.share_mask.data(), p.n_shared * sizeof(uint16_t));
    });
}
int bddb200_create_shard(const bddb200_instruction* instrs, size_t n_instr, const size_t* delims, size_t n_bdds, const double* costs_hi, size_t n_costs,
                         int precision, const bddb200_options* opts, int world, int rank, bddb200_shard_info* info, int32_t* new_of_old_out, bddb200_solver** out)
{
    (void)n_instr;
    if(out == nullptr || instrs == nullptr || delims == nullptr || info == nullptr || n_bdds == 0) { g_last_error = "null argument"; return BDDB200_ERR_INVALID_ARGUMENT; }
    *out = nullptr;
    bddb200_options o; bddb200_default_options(&o);
    if(opts) o = *opts;
    return guarded([&] {
        ShardPlan p;
        try { p = plan_shard(instrs, delims, n_bdds, std::max(n_costs, (size_t)o.nr_variables), world, rank); }
        catch(const std::invalid_argument& e) { throw api_error(BDDB200_ERR_INVALID_ARGUMENT, e.what()); }
        if(p.n_bdds == 0) throw api_error(BDDB200_ERR_INVALID_ARGUMENT, "create_shard: more ranks than BDDs (this rank's shard is empty)");
        info->nr_variables = p.n_vars; info->n_shared = p.n_shared; info->shared_entries = p.shared_entries; info->first_bdd = p.first_bdd; info->n_bdds = p.n_bdds;
        if(new_of_old_out) std::memcpy(new_of_old_out, p.new_of_old.data(), p.n_vars * sizeof(int32_t));
        // this rank's instructions with relabelled variables; child indices stay absolute (the base pointer is shifted instead)
        constexpr size_t BOT = (size_t)-2;
        const size_t i0 = delims[p.first_bdd], i1 = delims[p.first_bdd + p.n_bdds];
        std::vector<bddb200_instruction> local(instrs + i0, instrs + i1);
        for(bddb200_instruction& ins : local) if(ins.index < BOT) ins.index = (size_t)p.new_of_old[ins.index];
        std::vector<double> costs_new(p.n_vars, 0.0);
        for(size_t v = 0; v < std::min(n_costs, p.n_vars); ++v) costs_new[(size_t)p.new_of_old[v]] = costs_hi[v];
        o.nr_variables = p.n_vars; o.nr_bdds_per_var_host = p.counts_new.data(); o.n_shared_vars = p.n_shared;
        const bddb200_instruction* base = local.data() - i0;
        if(precision == BDDB200_DOUBLE) *out = new SolverImpl<double>(base, i1, delims + p.first_bdd, p.n_bdds, costs_new.data(), costs_new.size(), o);
        else if(precision == BDDB200_FLOAT) *out = new SolverImpl<float>(base, i1, delims + p.first_bdd, p.n_bdds, costs_new.data(), costs_new.size(), o);
        else throw api_error(BDDB200_ERR_INVALID_ARGUMENT, "precision must be BDDB200_FLOAT or BDDB200_DOUBLE");
        (*out)->set_push_masks(p.share_mask.data(), p.n_shared);
    });
}
void bddb200_destroy(bddb200_solver* s) { delete s; }
int bddb200_clone(const bddb200_solver* s, bddb200_solver** out)
{
    if(out == nullptr) { g_last_error = "null argument"; return BDDB200_ERR_INVALID_ARGUMENT; }
    *out = nullptr;
    REQUIRE_SOLVER(s);
    return guarded([&] { *out = s->clone(); });
}

size_t bddb200_nr_variables(const bddb200_solver* s) { return s ? s->nr_variables() : 0; }
size_t bddb200_nr_bdds(const bddb200_solver* s) { return s ? s->nr_bdds() : 0; }
size_t bddb200_nr_layers(const bddb200_solver* s) { return s ? s->nr_layers() : 0; }
size_t bddb200_nr_bdd_nodes(const bddb200_solver* s) { return s ? s->nr_bdd_nodes() : 0; }
size_t bddb200_nr_hops(const bddb200_solver* s) { return s ? s->nr_hops() : 0; }
int bddb200_precision_of(const bddb200_solver* s) { return s ? s->precision : -1; }
int bddb200_device_of(const bddb200_solver* s) { return s ? s->device : -1; }
int bddb200_nr_bdds_per_var(const bddb200_solver* s, int32_t* out) { REQUIRE_SOLVER(s); return guarded([&] { s->nr_bdds_per_var(out); }); }
int bddb200_layer_primal_indices(const bddb200_solver* s, int32_t* out) { REQUIRE_SOLVER(s); return guarded([&] { s->layer_primal_indices(out); }); }
int bddb200_layer_bdd_indices(const bddb200_solver* s, int32_t* out) { REQUIRE_SOLVER(s); return guarded([&] { s->layer_bdd_indices(out); }); }

int bddb200_iteration(bddb200_solver* s, double omega) { REQUIRE_SOLVER(s); return guarded([&] { s->iteration(omega); }); }
int bddb200_iterations(bddb200_solver* s, double omega, size_t n) { REQUIRE_SOLVER(s); return guarded([&] { s->iterations(omega, n); }); }
int bddb200_forward_pass(bddb200_solver* s, double omega) { REQUIRE_SOLVER(s); return guarded([&] { s->forward_pass(omega); }); }
int bddb200_backward_pass(bddb200_solver* s, double omega) { REQUIRE_SOLVER(s); return guarded([&] { s->backward_pass(omega); }); }
int bddb200_forward_mm(bddb200_solver* s, double omega, void* d) { REQUIRE_SOLVER(s); return guarded([&] { s->forward_mm(omega, d); }); }
int bddb200_backward_mm(bddb200_solver* s, double omega, void* d) { REQUIRE_SOLVER(s); return guarded([&] { s->backward_mm(omega, d); }); }
int bddb200_normalize_delta(const bddb200_solver* s, void* d) { REQUIRE_SOLVER(s); return guarded([&] { s->normalize_delta(d); }); }
int bddb200_get_delta(bddb200_solver* s, void* out, int out_is_host) { REQUIRE_SOLVER(s); return guarded([&] { s->get_delta(out, out_is_host); }); }
int bddb200_lower_bound(bddb200_solver* s, double* out) { REQUIRE_SOLVER(s); return guarded([&] { *out = s->lower_bound(); }); }
int bddb200_lower_bound_per_bdd(bddb200_solver* s, void* out) { REQUIRE_SOLVER(s); return guarded([&] { s->lower_bound_per_bdd(out); }); }

int bddb200_forward_run(bddb200_solver* s) { REQUIRE_SOLVER(s); return guarded([&] { s->forward_run(); }); }
int bddb200_backward_run(bddb200_solver* s) { REQUIRE_SOLVER(s); return guarded([&] { s->backward_run(); }); }
void bddb200_flush_forward_states(bddb200_solver* s) { if(s) s->flush_forward(); }
void bddb200_flush_backward_states(bddb200_solver* s) { if(s) s->flush_backward(); }

int bddb200_update_costs_host(bddb200_solver* s, const double* lo, size_t n_lo, const double* hi, size_t n_hi)
{ REQUIRE_SOLVER(s); return guarded([&] { s->update_costs_host(lo, n_lo, hi, n_hi); }); }
int bddb200_update_costs_host_real(bddb200_solver* s, const void* lo, size_t n_lo, const void* hi, size_t n_hi)
{ REQUIRE_SOLVER(s); return guarded([&] { s->update_costs_host_real(lo, n_lo, hi, n_hi); }); }
int bddb200_update_costs_dev(bddb200_solver* s, const void* lo, size_t n_lo, const void* hi, size_t n_hi)
{ REQUIRE_SOLVER(s); return guarded([&] { s->update_costs_dev(lo, n_lo, hi, n_hi); }); }
int bddb200_step_host(bddb200_solver* s, const void* lo, size_t n_lo, const void* hi, size_t n_hi, int src_is_real, double omega, double* lb_out)
{
    REQUIRE_SOLVER(s);
    if(lb_out == nullptr) { g_last_error = "null argument"; return BDDB200_ERR_INVALID_ARGUMENT; }
    return guarded([&] { *lb_out = s->step_host(lo, n_lo, hi, n_hi, src_is_real, omega); });
}
int bddb200_set_cost(bddb200_solver* s, double c, size_t var) { REQUIRE_SOLVER(s); return guarded([&] { s->set_cost(c, var); }); }
int bddb200_distribute_delta(bddb200_solver* s) { REQUIRE_SOLVER(s); return guarded([&] { s->distribute_delta(); }); }
int bddb200_get_solver_costs(const bddb200_solver* s, void* lo, void* hi, void* mmd) { REQUIRE_SOLVER(s); return guarded([&] { s->get_solver_costs(lo, hi, mmd); }); }
int bddb200_set_solver_costs(bddb200_solver* s, const void* lo, const void* hi, const void* mmd)
{
    REQUIRE_SOLVER(s);
    if(!lo || !hi || !mmd) { g_last_error = "null cost vector"; return BDDB200_ERR_INVALID_ARGUMENT; }
    return guarded([&] { s->set_solver_costs(lo, hi, mmd); });
}
int bddb200_primal_objective_host(bddb200_solver* s, double* out) { REQUIRE_SOLVER(s); return guarded([&] { s->primal_objective_host(out); }); }

int bddb200_min_marginals(bddb200_solver* s, int sorted, int32_t* primal, void* lo, void* hi)
{ REQUIRE_SOLVER(s); return guarded([&] { s->min_marginals(sorted, primal, lo, hi); }); }

int bddb200_min_marginals_host(bddb200_solver* s, int sorted, int32_t* primal, double* lo, double* hi)
{ REQUIRE_SOLVER(s); return guarded([&] { s->min_marginals_host(sorted, primal, lo, hi); }); }

int bddb200_bdds_solution(bddb200_solver* s, char* sol) { REQUIRE_SOLVER(s); return guarded([&] { s->bdds_solution(sol); }); }
int bddb200_net_solver_costs(const bddb200_solver* s, void* out) { REQUIRE_SOLVER(s); return guarded([&] { s->net_solver_costs(out); }); }
int bddb200_make_dual_feasible(const bddb200_solver* s, void* d) { REQUIRE_SOLVER(s); return guarded([&] { s->make_dual_feasible(d); }); }
int bddb200_gradient_step(bddb200_solver* s, const void* dir, double step) { REQUIRE_SOLVER(s); return guarded([&] { s->gradient_step(dir, step); }); }

int bddb200_synchronize(bddb200_solver* s) { REQUIRE_SOLVER(s); return guarded([&] { s->synchronize(); }); }
void* bddb200_stream(bddb200_solver* s) { return s ? s->stream_handle() : nullptr; }
size_t bddb200_kernel_launches(const bddb200_solver* s) { return s ? s->kernel_launches() : 0; }
int bddb200_delta_sum_buffer(bddb200_solver* s, void** out) { REQUIRE_SOLVER(s); return guarded([&] { *out = s->delta_sum_buffer(); }); }

// ---- primal rounding (incremental_mm_agreement_rounding_cuda.cu) ---------------------------------------------------------
int bddb200_rounding_perturb(bddb200_solver* s, double delta, int round_index, unsigned long long counts_out[4], char* types_dev, char* sol_host, int* solved)
{
    REQUIRE_SOLVER(s);
    if(counts_out == nullptr || solved == nullptr || !(delta > 0)) { g_last_error = "bddb200_rounding_perturb: invalid argument"; return BDDB200_ERR_INVALID_ARGUMENT; }
    return guarded([&] { *solved = s->rounding_perturb(delta, round_index, counts_out, types_dev, sol_host); });
}

int bddb200_run_solver(bddb200_solver* s, bddb200_lbfgs* lbfgs, size_t max_iter, double tolerance, double improvement_slope, double time_limit_s, double* lb_out)
{
    REQUIRE_SOLVER(s);
    return guarded([&] {
        if(lbfgs) run_solver_native(s, [&] { lbfgs->iteration(); }, max_iter, tolerance, improvement_slope, time_limit_s);
        else run_solver_native(s, [&] { s->iteration(0.5); }, max_iter, tolerance, improvement_slope, time_limit_s);
        if(lb_out) *lb_out = s->lower_bound();
    });
}

int bddb200_incremental_mm_agreement_rounding(bddb200_solver* s, bddb200_lbfgs* lbfgs, double init_delta, double delta_growth_rate, int num_itr_lb, int num_rounds,
                                              char* sol_host, int* solved, int* rounds_used)
{
    REQUIRE_SOLVER(s);
    if(!(init_delta > 0) || !(delta_growth_rate > 0) || sol_host == nullptr || solved == nullptr)
    { g_last_error = "bddb200_incremental_mm_agreement_rounding: invalid argument"; return BDDB200_ERR_INVALID_ARGUMENT; }
    return guarded([&] {
        *solved = 0;
        s->distribute_delta();                                        // :344
        double cur_delta = init_delta / delta_growth_rate;
        int round = 0;
        for(; round < num_rounds; ++round)
        {
            cur_delta = std::min(cur_delta * delta_growth_rate, 1e6);
            unsigned long long counts[4];
            if(lbfgs) lbfgs->flush();                                 // lbfgs<>::update_costs flushes the history (lbfgs_impl.h:343-348)
            if(s->rounding_perturb(cur_delta, round, counts, nullptr, sol_host)) { *solved = 1; break; }
            if(lbfgs) run_solver_native(s, [&] { lbfgs->iteration(); }, (size_t)num_itr_lb, 1e-7, 0.0001, std::numeric_limits<double>::max());
            else run_solver_native(s, [&] { s->iteration(0.5); }, (size_t)num_itr_lb, 1e-7, 0.0001, std::numeric_limits<double>::max());     // :367
        }
        if(rounds_used) *rounds_used = round + (*solved ? 1 : 0);
    });
}

// ---- L-BFGS ("lbfgs cuda mma"), lbfgs.cuh ------------------------------------------------------------------------
int bddb200_lbfgs_create(bddb200_solver* s, int history_size, double init_step_size, double req_rel_lb_increase,
                         double step_size_decrease_factor, double step_size_increase_factor, bddb200_lbfgs** out)
{
    if(out == nullptr) { g_last_error = "null argument"; return BDDB200_ERR_INVALID_ARGUMENT; }
    *out = nullptr;
    REQUIRE_SOLVER(s);
    LbfgsOptions o;
    if(history_size != 0) o.history_size = history_size;
    if(init_step_size > 0) o.init_step_size = init_step_size;
    if(req_rel_lb_increase > 0) o.req_rel_lb_increase = req_rel_lb_increase;
    if(step_size_decrease_factor > 0) o.step_size_decrease_factor = step_size_decrease_factor;
    if(step_size_increase_factor > 0) o.step_size_increase_factor = step_size_increase_factor;
    if(o.history_size < 2 || o.history_size > 64) { g_last_error = "lbfgs history size must be in [2, 64]"; return BDDB200_ERR_INVALID_ARGUMENT; }   // assert(m > 1), lbfgs_impl.h:30
    return guarded([&] {
        CUDA_CHECK(cudaSetDevice(s->device));
        if(s->precision == BDDB200_DOUBLE) *out = new LbfgsImpl<double>(s, o); else *out = new LbfgsImpl<float>(s, o);
    });
}
void bddb200_lbfgs_destroy(bddb200_lbfgs* l) { delete l; }
int bddb200_lbfgs_iteration(bddb200_lbfgs* l)
{
    if(l == nullptr) { g_last_error = "null lbfgs handle"; return BDDB200_ERR_INVALID_ARGUMENT; }
    return guarded([&] { l->iteration(); });
}
int bddb200_lbfgs_flush(bddb200_lbfgs* l)
{
    if(l == nullptr) { g_last_error = "null lbfgs handle"; return BDDB200_ERR_INVALID_ARGUMENT; }
    return guarded([&] { l->flush(); });
}
int bddb200_lbfgs_stats(const bddb200_lbfgs* l, size_t* lbfgs_iterations, size_t* mma_iterations, double* step_size)
{
    if(l == nullptr) { g_last_error = "null lbfgs handle"; return BDDB200_ERR_INVALID_ARGUMENT; }
    if(lbfgs_iterations) *lbfgs_iterations = l->lbfgs_iterations();
    if(mma_iterations) *mma_iterations = l->mma_iterations();
    if(step_size) *step_size = l->step_size();
    return BDDB200_OK;
}

int bddb200_save_size(const bddb200_solver* s, size_t* bytes_out)
{
    REQUIRE_SOLVER(s);
    return guarded([&] { std::vector<unsigned char> b; s->save(b); *bytes_out = b.size(); });
}
int bddb200_save(const bddb200_solver* s, void* buf, size_t bytes, size_t* written_out)
{
    REQUIRE_SOLVER(s);
    return guarded([&] {
        std::vector<unsigned char> b; s->save(b);
        if(written_out) *written_out = b.size();
        if(buf == nullptr || bytes < b.size()) throw api_error(BDDB200_ERR_INVALID_ARGUMENT, "bddb200_save: buffer too small (ask bddb200_save_size)");
        std::memcpy(buf, b.data(), b.size());
    });
}
int bddb200_load(const void* buf, size_t bytes, int device, bddb200_solver** out)
{
    if(buf == nullptr || out == nullptr || bytes < 12) { g_last_error = "bddb200_load: invalid argument"; return BDDB200_ERR_INVALID_ARGUMENT; }
    return guarded([&] {
        const unsigned char* p = static_cast<const unsigned char*>(buf);
        uint64_t magic; int32_t prec;
        std::memcpy(&magic, p, 8); std::memcpy(&prec, p + 8, 4);
        if(magic != 0x3130533030324442ull || (prec != BDDB200_FLOAT && prec != BDDB200_DOUBLE)) throw api_error(BDDB200_ERR_INVALID_ARGUMENT, "bddb200_load: not a solver blob of this library version");
        if(prec == BDDB200_DOUBLE) *out = new SolverImpl<double>(p + 12, bytes - 12, device);
        else *out = new SolverImpl<float>(p + 12, bytes - 12, device);
    });
}
int bddb200_delta_sum_index(const bddb200_solver* s, int* out) { REQUIRE_SOLVER(s); return guarded([&] { *out = s->delta_sum_index(); }); }
int bddb200_set_push_masks(bddb200_solver* s, const uint16_t* masks_host, size_t n)
{
    REQUIRE_SOLVER(s);
    if(masks_host == nullptr && n > 0) { g_last_error = "null argument"; return BDDB200_ERR_INVALID_ARGUMENT; }
    return guarded([&] { s->set_push_masks(masks_host, n); });
}
int bddb200_push_exchange_supported(const bddb200_solver* s, int* out) { REQUIRE_SOLVER(s); return guarded([&] { *out = s->push_exchange_supported(); }); }
int bddb200_set_delta_buffers(bddb200_solver* s, void* b0, void* b1, void* b2)
{
    REQUIRE_SOLVER(s);
    if((b0 == nullptr) != (b1 == nullptr) || (b0 == nullptr) != (b2 == nullptr)) { g_last_error = "give three buffers or none"; return BDDB200_ERR_INVALID_ARGUMENT; }
    return guarded([&] { s->set_delta_buffers(b0, b1, b2); });
}
int bddb200_set_delta_input(bddb200_solver* s, void* in, size_t n_shared_vars) { REQUIRE_SOLVER(s); return guarded([&] { s->set_delta_input(in, n_shared_vars); }); }

int bddb200_set_exchange(bddb200_solver* s, int world, int rank, const void* const* peer_bufs_dev, uint32_t* const* flags_dev, void* out_dev,
                         void* const* peer_outs_dev, const void* mc_in, void* mc_out, size_t n_exchange, int mode)
{ REQUIRE_SOLVER(s); return guarded([&] { s->set_exchange(world, rank, peer_bufs_dev, flags_dev, out_dev, peer_outs_dev, mc_in, mc_out, n_exchange, mode); }); }

int bddb200_delta_exchange(void* stream, int precision, int world, int rank, const void* const* peer_bufs_dev, uint32_t* const* flags_dev,
                           uint32_t epoch, size_t offset_elems, void* out_dev, size_t n_exchange)
{
    if(world < 1 || world > EXCHANGE_MAX_WORLD || rank < 0 || rank >= world || peer_bufs_dev == nullptr || flags_dev == nullptr || out_dev == nullptr
       || (n_exchange & 1))
    { g_last_error = "bddb200_delta_exchange: invalid argument"; return BDDB200_ERR_INVALID_ARGUMENT; }
    return guarded([&] {
        int dev = 0, sms = 0;
        CUDA_CHECK(cudaGetDevice(&dev));
        CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const size_t pairs = n_exchange / 2;
        const unsigned blocks = (unsigned)std::max<size_t>(1, std::min<size_t>((pairs + 255) / 256, (size_t)sms * 4));
        cudaStream_t st = (cudaStream_t)stream;
        if(precision == BDDB200_DOUBLE)
            delta_exchange_kernel<double><<<blocks, 256, 0, st>>>(reinterpret_cast<const double* const*>(peer_bufs_dev), flags_dev, world, rank, epoch, offset_elems,
                                                                  static_cast<double*>(out_dev), pairs, nullptr);
        else
            delta_exchange_kernel<float><<<blocks, 256, 0, st>>>(reinterpret_cast<const float* const*>(peer_bufs_dev), flags_dev, world, rank, epoch, offset_elems,
                                                                 static_cast<float*>(out_dev), pairs, nullptr);
        CUDA_CHECK(cudaGetLastError());
    });
}

int bddb200_delta_exchange_two_shot(void* stream, int precision, int world, int rank, const void* const* peer_bufs_dev, void* const* peer_outs_dev,
                                    uint32_t* const* flags_dev, uint32_t epoch, size_t offset_elems, size_t n_exchange)
{
    if(world < 2 || world > EXCHANGE_MAX_WORLD || rank < 0 || rank >= world || peer_bufs_dev == nullptr || peer_outs_dev == nullptr || flags_dev == nullptr || (n_exchange & 1))
    { g_last_error = "bddb200_delta_exchange_two_shot: invalid argument"; return BDDB200_ERR_INVALID_ARGUMENT; }
    return guarded([&] {
        int dev = 0, sms = 0;
        CUDA_CHECK(cudaGetDevice(&dev));
        CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const size_t pairs = n_exchange / 2, per = (pairs + world - 1) / world;
        const unsigned blocks = (unsigned)std::max<size_t>(1, std::min<size_t>((per + 255) / 256, (size_t)sms * 4));     // co-resident: the CTAs wait for each other
        cudaStream_t st = (cudaStream_t)stream;
        if(precision == BDDB200_DOUBLE)
            delta_exchange2_kernel<double><<<blocks, 256, 0, st>>>(reinterpret_cast<const double* const*>(peer_bufs_dev), reinterpret_cast<double* const*>(peer_outs_dev),
                                                                   flags_dev, world, rank, epoch, offset_elems, pairs, nullptr);
        else
            delta_exchange2_kernel<float><<<blocks, 256, 0, st>>>(reinterpret_cast<const float* const*>(peer_bufs_dev), reinterpret_cast<float* const*>(peer_outs_dev),
                                                                  flags_dev, world, rank, epoch, offset_elems, pairs, nullptr);
        CUDA_CHECK(cudaGetLastError());
    });
}

int bddb200_trace_pass(bddb200_solver* s, int forward, double omega, unsigned long long* out_host, size_t max_bundles, size_t* n_out)
{ REQUIRE_SOLVER(s); return guarded([&] { *n_out = s->trace_pass(forward, omega, out_host, max_bundles); }); }

int bddb200_layout_stats(const bddb200_instruction* instrs, size_t n_instr, const size_t* delims, size_t n_bdds, int lanes_per_bdd, uint64_t* out, size_t n)
{
    return guarded([&] {
        const HostLayout L = build_layout(instrs, n_instr, delims, n_bdds, lanes_per_bdd);
        const uint64_t vals[13] = {L.n_slots, L.n_lay, L.bundles.size(), L.n_real_nodes, L.max_hops,
                                   std::max(L.max_tile_small, L.max_tile_large), L.n_small_bundles, L.n_layers_ext,
                                   L.chunks.size(), L.stage_small, L.stage_large, L.n_lane_bundles, L.n_topo};
        for(size_t i = 0; i < n && i < 13; ++i) out[i] = vals[i];
    });
}

} // extern "C"
