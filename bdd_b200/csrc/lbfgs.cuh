// bdd_b200/csrc/lbfgs.cuh -- L-BFGS acceleration of the deferred MMA solver, on the device.
//
// Reference: template lbfgs<SOLVER, VECTOR, REAL, INT_VECTOR, CUDA_SOLVER> (include/bdd_solver/lbfgs.h:35-110,
// src/bdd_solver/lbfgs_impl.h:46-420) wrapped around bdd_cuda_parallel_mma<REAL> ("lbfgs cuda mma",
// src/bdd_solver/bdd_solver.cpp:222-236).  At the reference commit every CUDA branch of that template is compiled
// out (`#ifdef CUDACC`, never defined) and the first loop's alpha is shadowed (lbfgs_impl.h:251-263), so the
// reference L-BFGS does not work on either back end: parity for this file is UNPINNED.  What is
// implemented here is the algorithm that code spells out, with those two defects removed:
//   iteration()                      lbfgs_impl.h:138-157
//   store_iterate                    :46-135   s = x - x_prev, y = g_prev - g, keep if <s, y> > 1e-8, history <= m
//   compute_update_direction         :226-316  two-loop recursion, H0 scaling folded into the oldest pair (:286-292)
//   search_step_size_and_apply       :159-224  <= 7 trial steps judged by the lower bound
//   lbfgs_update_possible            :335-341  history full and <= 5 consecutive failures
// x = net_solver_costs() (hi - lo + deferred mm difference per layer), g = the per-BDD argmin solution (a subgradient
// of the lower bound), vectors have nr_layers() entries in layer order, terminal layers 0.
//
// B200 shape of it: all vectors stay on the device; one fused kernel per history pair in each loop of the
// recursion (axpy with the previous coefficient + the dot product that yields the next one, accumulated in double in a
// fixed order, so a solve is bit-reproducible),
// coefficients are passed from kernel to kernel through device memory, so computing a direction is 2m + 1 launches and
// no host synchronisation.  The only host round trips are the curvature test (one scalar per iteration) and the lower
// bounds of the step-size search, which are decisions the host loop takes.
#pragma once

#include <cuda_runtime.h>
#include <deque>
#include <memory>
#include <vector>

namespace bddb200 {

constexpr int LBFGS_THREADS = 256;

// Grid-wide sum in a FIXED order (bit-reproducible run to run: the step-size search compares lower bounds, and a last-bit
// difference in a coefficient can flip one of its decisions): every block writes its partial sum, the last block to arrive adds
// the partials in index order and stores the total.  `counter` must be zero on entry and is left zero.
__device__ __forceinline__ void grid_sum_to(double v, double* target, double* partials, unsigned int* counter)
{
    __shared__ double sh[LBFGS_THREADS / 32];
    __shared__ bool last;
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();                        // sh may still be read by a previous call
    if(lane == 0) sh[warp] = v;
    __syncthreads();
    if(threadIdx.x == 0)
    {
        double t = 0;
        for(int w = 0; w < LBFGS_THREADS / 32; ++w) t += sh[w];
        partials[blockIdx.x] = t;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if(last)
    {   // the whole last block adds the partials: thread t takes partials t, t + 256, ... in that order, then a fixed-shape tree
        // (a single thread walking ~1 200 partials cost 35 us per call -- eleven calls per L-BFGS direction)
        __shared__ double tree[LBFGS_THREADS];
        __threadfence();
        double t = 0;
        for(unsigned int b = threadIdx.x; b < gridDim.x; b += LBFGS_THREADS) t += __ldcg(partials + b);
        tree[threadIdx.x] = t;
        __syncthreads();
        for(int o = LBFGS_THREADS / 2; o > 0; o >>= 1)
        {
            if((int)threadIdx.x < o) tree[threadIdx.x] += tree[threadIdx.x + o];
            __syncthreads();
        }
        if(threadIdx.x == 0) { *target = tree[0]; *counter = 0; }
    }
}

// store_iterate: s = x - x_prev, y = g_prev - g (the problem is a maximisation, lbfgs_impl.h:83), out[0] += <s, y>,
// out[1] += <y, y>; x_prev = x, g_prev = g.
template<typename REAL>
__global__ void __launch_bounds__(LBFGS_THREADS) lbfgs_store_kernel(const REAL* __restrict__ x, const char* __restrict__ g, REAL* __restrict__ x_prev, char* __restrict__ g_prev,
                                                                    REAL* __restrict__ s, signed char* __restrict__ y, double* __restrict__ out, double* __restrict__ partials,
                                                                    unsigned int* __restrict__ counters, size_t n, int have_prev)
{
    double sy = 0, yy = 0;
    for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        const REAL xi = x[i];
        const char gi = g[i];
        if(have_prev)
        {
            const REAL si = xi - x_prev[i];
            const signed char yi = (signed char)(g_prev[i] - gi);
            s[i] = si; y[i] = yi;
            sy += (double)si * (double)yi; yy += (double)yi * (double)yi;
        }
        x_prev[i] = xi; g_prev[i] = gi;
    }
    if(have_prev) { grid_sum_to(sy, out, partials, counters); grid_sum_to(yy, out + 1, partials + gridDim.x, counters + 1); }
}

// One step of the two-loop recursion:  d += coef * v  (v = y pair or s pair of the previous step; coef is computed from
// device scalars), then  out += <w, d>  (w = the vector whose coefficient the next step needs).
//   mode 0: d = g (start), no update
//   mode 1: first loop,  coef = -alpha_j,             alpha_j = dots_a[j] / rho_inv_j,                 v = y_j
//   mode 2: second loop, coef = alpha_j - beta_j,      beta_j = rho_j' * dots_b[j],                     v = s_j
template<typename REAL>
__global__ void __launch_bounds__(LBFGS_THREADS) lbfgs_step_kernel(REAL* __restrict__ d, const char* __restrict__ g, int mode,
                                                                   const REAL* __restrict__ v_s, const signed char* __restrict__ v_y,
                                                                   const double* __restrict__ alpha_dot, double rho_inv, const double* __restrict__ beta_dot, double rho_scaled,
                                                                   const REAL* __restrict__ w_s, const signed char* __restrict__ w_y, double* __restrict__ out,
                                                                   double* __restrict__ partials, unsigned int* __restrict__ counters, size_t n)
{
    double coef = 0;
    if(mode == 1) coef = -(*alpha_dot / rho_inv);
    else if(mode == 2) coef = *alpha_dot / rho_inv - rho_scaled * *beta_dot;
    double acc = 0;
    for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        REAL di;
        if(mode == 0) di = (REAL)g[i];
        else
        {
            const REAL vi = v_s ? v_s[i] : (REAL)v_y[i];
            di = d[i] + (REAL)coef * vi;
        }
        d[i] = di;
        if(out) acc += (double)di * (double)(w_s ? w_s[i] : (REAL)w_y[i]);
    }
    if(out) grid_sum_to(acc, out, partials, counters);
}

struct LbfgsOptions {
    int history_size = 5;                       // lbfgs_default_history_size, lbfgs.h:29-33
    double init_step_size = 1e-6;
    double req_rel_lb_increase = 1e-6;
    double step_size_decrease_factor = 0.8;
    double step_size_increase_factor = 1.1;
};

} // namespace bddb200

struct bddb200_lbfgs {
    virtual ~bddb200_lbfgs() {}
    virtual void iteration() = 0;
    virtual void flush() = 0;
    virtual size_t lbfgs_iterations() const = 0;
    virtual size_t mma_iterations() const = 0;
    virtual double step_size() const = 0;
};

namespace bddb200 {

// (DevBuf<T> is the owning device buffer of bdd_b200.cu, which includes this file)
template<typename REAL>
class LbfgsImpl final : public bddb200_lbfgs {
public:
    LbfgsImpl(bddb200_solver* solver, const LbfgsOptions& o) : s_(solver), opt_(o), m_(o.history_size), step_size_(o.init_step_size)
    {
        n_ = s_->nr_layers();
        stream_ = (cudaStream_t)s_->stream_handle();
        x_.alloc(n_); g_.alloc(n_); x_prev_.alloc(n_); g_prev_.alloc(n_); d_.alloc(n_);
        s_new_.alloc(n_); y_new_.alloc(n_);
        hist_s_.reset(new DevBuf<REAL>[m_]); hist_y_.reset(new DevBuf<signed char>[m_]);
        for(int i = 0; i < m_; ++i) { hist_s_[i].alloc(n_); hist_y_[i].alloc(n_); free_slots_.push_back(i); }
        scal_.alloc(2 + 2 * (size_t)m_);
        partials_.alloc(2 * (size_t)blocks_for_n()); counters_.alloc(2);
        cudaMemsetAsync(counters_.p, 0, 2 * sizeof(unsigned int), stream_);
        if(cudaMallocHost(&h_scal_, 2 * sizeof(double)) != cudaSuccess) throw std::runtime_error("cudaMallocHost failed");
        blocks_ = blocks_for_n();
    }
    unsigned blocks_for_n() const
    {
        int dev = 0, sms = 0;
        cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        return (unsigned)std::max<size_t>(1, std::min<size_t>((s_->nr_layers() + LBFGS_THREADS - 1) / LBFGS_THREADS, (size_t)sms * 8));
    }
    ~LbfgsImpl() override { if(h_scal_) cudaFreeHost(h_scal_); }

    size_t lbfgs_iterations() const override { return n_lbfgs_; }
    size_t mma_iterations() const override { return n_mma_; }
    double step_size() const override { return step_size_; }

    // flush_lbfgs_states, lbfgs_impl.h:318-327 (called when the costs change from outside)
    void flush() override
    {
        unsuccessful_ = 0;
        for(const Pair& p : history_) free_slots_.push_back(p.slot);
        history_.clear();
        prev_stored_ = false;
    }

    // lbfgs_impl.h:138-157
    void iteration() override
    {
        if(lb_history_.empty()) lb_history_.push_back(s_->lower_bound());
        s_->bdds_solution(g_.p);                       // bdds_solution_vec: per-BDD argmin, a subgradient of the bound
        store_iterate();
        if((int)history_.size() >= m_ && unsuccessful_ <= 5) lbfgs_iteration();        // choose_solver, :409-418
        else { s_->iteration(0.5); ++n_mma_; }
        lb_history_.push_back(s_->lower_bound());
        if(lb_history_.size() > (size_t)(m_ + 8)) lb_history_.pop_front();
    }

private:
    struct Pair { int slot; double rho_inv, yy; };

    static void check(cudaError_t e, const char* what) { if(e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e)); }

    // lbfgs_impl.h:46-135
    void store_iterate()
    {
        s_->net_solver_costs(x_.p);
        check(cudaMemsetAsync(scal_.p, 0, 2 * sizeof(double), stream_), "memset");
        lbfgs_store_kernel<REAL><<<blocks_, LBFGS_THREADS, 0, stream_>>>(x_.p, g_.p, x_prev_.p, g_prev_.p, s_new_.p, y_new_.p, scal_.p, partials_.p, counters_.p, n_, prev_stored_ ? 1 : 0);
        check(cudaGetLastError(), "lbfgs_store_kernel");
        if(!prev_stored_) { prev_stored_ = true; return; }
        check(cudaMemcpyAsync(h_scal_, scal_.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream_), "memcpy");
        check(cudaStreamSynchronize(stream_), "sync");
        const double rho_inv = h_scal_[0], yy = h_scal_[1];
        if(rho_inv > 1e-8)                            // otherwise the curvature condition is not strongly satisfied: skip the pair
        {
            if((int)history_.size() == m_) { free_slots_.push_back(history_.front().slot); history_.pop_front(); }
            const int slot = free_slots_.back(); free_slots_.pop_back();
            std::swap(hist_s_[slot].p, s_new_.p);        // the new pair moves into the history without a copy
            std::swap(hist_y_[slot].p, y_new_.p);
            history_.push_back(Pair{slot, rho_inv, yy});
        }
        else prev_stored_ = false;                     // (x_prev / g_prev were still updated, as in the reference)
    }

    // lbfgs_impl.h:226-316, with alpha actually carried from the first loop to the second
    void compute_update_direction()
    {
        const int k = (int)history_.size();
        double* dots_a = scal_.p + 2;                  // <s_i, q> of the first loop
        double* dots_b = scal_.p + 2 + m_;             // <y_i, r> of the second loop
        check(cudaMemsetAsync(dots_a, 0, 2 * (size_t)m_ * sizeof(double), stream_), "memset");
        auto S = [&](int i) { return hist_s_[history_[i].slot].p; };
        auto Y = [&](int i) { return hist_y_[history_[i].slot].p; };
        // q = g;  dots_a[k-1] = <s_{k-1}, q>
        lbfgs_step_kernel<REAL><<<blocks_, LBFGS_THREADS, 0, stream_>>>(d_.p, g_.p, 0, nullptr, nullptr, nullptr, 1.0, nullptr, 0.0, S(k - 1), nullptr, dots_a + (k - 1), partials_.p, counters_.p, n_);
        // first loop, newest to oldest: q -= alpha_i y_i; the same launch takes the dot product the next step needs
        for(int i = k - 1; i >= 0; --i)
        {
            const bool last = i == 0;
            lbfgs_step_kernel<REAL><<<blocks_, LBFGS_THREADS, 0, stream_>>>(d_.p, nullptr, 1, nullptr, Y(i), dots_a + i, history_[i].rho_inv, nullptr, 0.0,
                                                                           last ? nullptr : S(i - 1), last ? Y(0) : nullptr, last ? dots_b + 0 : dots_a + (i - 1), partials_.p, counters_.p, n_);
        }
        // second loop, oldest to newest: r += (alpha_i - beta_i) s_i, beta_i = rho_i <y_i, r>; the initial Hessian scaling
        // rho_inv_last / <y_last, y_last> multiplies the oldest pair's rho (lbfgs_impl.h:286-292)
        const double h0 = history_.back().rho_inv / (1e-8 + history_.back().yy);
        for(int i = 0; i < k; ++i)
        {
            const bool last = i == k - 1;
            double rho = 1.0 / history_[i].rho_inv;
            if(i == 0) rho *= h0;
            lbfgs_step_kernel<REAL><<<blocks_, LBFGS_THREADS, 0, stream_>>>(d_.p, nullptr, 2, S(i), nullptr, dots_a + i, history_[i].rho_inv, dots_b + i, rho,
                                                                           nullptr, last ? nullptr : Y(i + 1), last ? nullptr : dots_b + (i + 1), partials_.p, counters_.p, n_);
        }
        check(cudaGetLastError(), "lbfgs_step_kernel");
    }

    // lbfgs_impl.h:159-224
    void search_step_size_and_apply()
    {
        const double lb_pre = s_->lower_bound();
        const double past_lb_increase = *(lb_history_.rbegin() + (m_ - 2)) - *(lb_history_.rbegin() + (m_ - 1));
        auto rel_change = [&]() { return (s_->lower_bound() - lb_pre) / (1e-9 + past_lb_increase); };
        double prev_step = 0.0;
        auto apply = [&](double new_step) {
            const double net = new_step - prev_step;
            if(net != 0.0) s_->gradient_step(d_.p, net);
            prev_step = new_step;
        };
        size_t num_updates = 0;
        double cur = 0.0, best_step = 0.0, best_rel = 0.0;
        do
        {
            apply(step_size_);
            cur = rel_change();
            if(best_rel < cur) { best_rel = cur; best_step = step_size_; }
            if(cur <= 0.0) step_size_ *= opt_.step_size_decrease_factor;
            else if(cur < opt_.req_rel_lb_increase) step_size_ *= opt_.step_size_increase_factor;
            if(num_updates > 5)
            {
                if(best_rel > opt_.req_rel_lb_increase / 10.0) apply(best_step);
                else { apply(0.0); ++unsuccessful_; }
                return;
            }
            ++num_updates;
        } while(cur < opt_.req_rel_lb_increase);
        if(num_updates == 1 && unsuccessful_ == 0) step_size_ *= opt_.step_size_increase_factor;
        unsuccessful_ = 0;
    }

    // lbfgs_impl.h:384-406
    void lbfgs_iteration()
    {
        compute_update_direction();
        s_->make_dual_feasible(d_.p);          // per-variable mean removal: the step keeps the costs a valid reparametrisation
        search_step_size_and_apply();
        s_->iteration(0.5);
        ++n_lbfgs_;
    }

    bddb200_solver* s_;
    LbfgsOptions opt_;
    int m_;
    double step_size_;
    size_t n_ = 0;
    cudaStream_t stream_ = nullptr;
    unsigned blocks_ = 1;
    DevBuf<REAL> x_, x_prev_, d_, s_new_;
    DevBuf<char> g_, g_prev_;
    DevBuf<signed char> y_new_;
    std::unique_ptr<DevBuf<REAL>[]> hist_s_;
    std::unique_ptr<DevBuf<signed char>[]> hist_y_;
    std::vector<int> free_slots_;
    std::deque<Pair> history_;
    DevBuf<double> scal_, partials_;
    DevBuf<unsigned int> counters_;
    double* h_scal_ = nullptr;
    std::deque<double> lb_history_;
    bool prev_stored_ = false;
    int unsuccessful_ = 0;
    size_t n_lbfgs_ = 0, n_mma_ = 0;
};

} // namespace bddb200
