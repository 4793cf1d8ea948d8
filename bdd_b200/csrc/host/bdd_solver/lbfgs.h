// bdd_b200/csrc/host/bdd_solver/lbfgs.h -- drop-in for the reference's include/bdd_solver/lbfgs.h.
//
// It takes the place of the reference's header (which moves to bdd_solver/lbfgs_generic.h) or sits ahead of it on the
// include path.  The reference header is pulled in unchanged, so lbfgs<SOLVER, ...> over the CPU solvers is the reference's own;
// for the GPU solver the partial specialisation below replaces it:
//     lbfgs<bdd_cuda_parallel_mma<REAL>, VECTOR, REAL, INT_VECTOR, true>
// (bdd_solver.h:60-61: cuda_lbfgs_parallel_mma_{float,double}_type) -> bddb200_lbfgs_* of libbdd_b200.so, where the history, the
// two-loop recursion and the step-size search run on the device (bdd_b200/csrc/lbfgs.cuh).  The reference's generic template is
// dead code for CUDA solvers at this commit (its CUDA branches are guarded by `#ifdef CUDACC`).
// Same constructors, iteration(), update_costs(); copyable / movable like the reference class (a copy starts with an empty history).
#pragma once

#if __has_include("bdd_solver/lbfgs_generic.h")
#include "bdd_solver/lbfgs_generic.h"        // the reference's lbfgs.h under this name, when this file has taken its place in the tree
#else
#include_next "bdd_solver/lbfgs.h"           // this file merely sits ahead of the reference's include directory (tests/cpp/test_dropin.cu)
#endif

#include "bdd_solver/bdd_cuda_parallel_mma.h"

namespace LPMP {

template<typename REAL, typename VECTOR, typename INT_VECTOR>
class lbfgs<bdd_cuda_parallel_mma<REAL>, VECTOR, REAL, INT_VECTOR, true> : public bdd_cuda_parallel_mma<REAL>
{
    using SOLVER = bdd_cuda_parallel_mma<REAL>;
    public:
        lbfgs() {}
        lbfgs(const BDD::bdd_collection& bdd_col, const std::vector<double>& costs_hi,
              const int history_size = lbfgs_default_history_size, const double init_step_size = lbfgs_default_init_step_size,
              const double req_rel_lb_increase = lbfgs_default_req_rel_lb_increase,
              const double step_size_decrease_factor = lbfgs_default_step_size_decrease_factor,
              const double step_size_increase_factor = lbfgs_default_step_size_increase_factor)
            : SOLVER(bdd_col, costs_hi), m_(history_size), step_(init_step_size), req_(req_rel_lb_increase), dec_(step_size_decrease_factor), inc_(step_size_increase_factor)
        { create(); }
        lbfgs(const BDD::bdd_collection& bdd_col,
              const int history_size = lbfgs_default_history_size, const double init_step_size = lbfgs_default_init_step_size,
              const double req_rel_lb_increase = lbfgs_default_req_rel_lb_increase,
              const double step_size_decrease_factor = lbfgs_default_step_size_decrease_factor,
              const double step_size_increase_factor = lbfgs_default_step_size_increase_factor)
            : lbfgs(bdd_col, std::vector<double>{}, history_size, init_step_size, req_rel_lb_increase, step_size_decrease_factor, step_size_increase_factor) {}

        lbfgs(const lbfgs& o) : SOLVER(o), m_(o.m_), step_(o.step_), req_(o.req_), dec_(o.dec_), inc_(o.inc_) { create(); }
        lbfgs(lbfgs&& o) noexcept : SOLVER(std::move(o)), l_(o.l_), m_(o.m_), step_(o.step_), req_(o.req_), dec_(o.dec_), inc_(o.inc_) { o.l_ = nullptr; }
        lbfgs& operator=(const lbfgs& o) { if(this != &o) { destroy(); SOLVER::operator=(o); copy_params(o); create(); } return *this; }
        lbfgs& operator=(lbfgs&& o) noexcept { if(this != &o) { destroy(); SOLVER::operator=(std::move(o)); copy_params(o); l_ = o.l_; o.l_ = nullptr; } return *this; }
        ~lbfgs() { destroy(); }

        void iteration() { bddb200_detail::check(bddb200_lbfgs_iteration(l_)); }

        void update_costs(const std::vector<REAL>& cost_lo, const std::vector<REAL>& cost_hi)
        {
            bddb200_detail::check(bddb200_lbfgs_flush(l_));          // flush_lbfgs_states, lbfgs_impl.h:343-348
            SOLVER::update_costs(cost_lo, cost_hi);
        }
        void update_costs(const thrust::device_vector<REAL>& cost_0, const thrust::device_vector<REAL>& cost_1)
        {
            bddb200_detail::check(bddb200_lbfgs_flush(l_));
            SOLVER::update_costs(cost_0, cost_1);
        }

    private:
        void create() { if(this->h_) bddb200_detail::check(bddb200_lbfgs_create(this->h_, m_, step_, req_, dec_, inc_, &l_)); }
        void destroy() { if(l_) { bddb200_lbfgs_destroy(l_); l_ = nullptr; } }
        void copy_params(const lbfgs& o) { m_ = o.m_; step_ = o.step_; req_ = o.req_; dec_ = o.dec_; inc_ = o.inc_; }
        bddb200_lbfgs* l_ = nullptr;
        int m_ = lbfgs_default_history_size;
        double step_ = lbfgs_default_init_step_size, req_ = lbfgs_default_req_rel_lb_increase;
        double dec_ = lbfgs_default_step_size_decrease_factor, inc_ = lbfgs_default_step_size_increase_factor;
};

}
