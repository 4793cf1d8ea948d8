// bdd_b200/csrc/host/bdd_solver/bdd_cuda_parallel_mma.h -- drop-in replacement of the reference's
//   include/bdd_solver/bdd_cuda_parallel_mma.h  (class LPMP::bdd_cuda_parallel_mma<REAL>, :7-52)
//   include/bdd_solver/bdd_cuda_base.h          (class LPMP::bdd_cuda_base<REAL>, :57-226)
// on top of the C ABI of libbdd_b200.so (include/bdd_b200.h).
//
// Put this file and bdd_cuda_base.h in the place of the reference's two headers (include/bdd_solver/bdd_solver.h includes them by
// bare name, so being ahead on the include path is not enough for it) and link libbdd_b200.so instead of
// compiling src/bdd_solver/bdd_cuda_base.cu and bdd_cuda_parallel_mma.cu: every caller of the class --
// run_solver (include/run_solver_util.h:27-76), lbfgs<> (include/bdd_solver/lbfgs.h:22-27), the JSON driver
// (src/bdd_solver/bdd_solver.cpp), the GPU rounding (src/bdd_solver/incremental_mm_agreement_rounding_cuda.cu:264-372), the pimpl
// bdd_cuda<REAL> (src/bdd_cuda.cpp), the hybrid solver (src/bdd_solver/bdd_multi_parallel_mma_base.cu:130-157) and the reference's
// own tests of the class -- compiles unchanged (`make -C oracle integration` does exactly that): same class name, same method
// names, argument types (thrust::device_vector like the reference) and error behaviour (std::runtime_error).
//
// Host-side glue only: no kernel lives here.  Needs nvcc (thrust::device_vector) and the
// reference's bdd_collection / two_dimensional_variable_array headers, like the file it replaces.
//
// Differences to the reference class, all outside what its callers use:
//  * the per-layer vectors are in BDD-major "layer order" (include/bdd_b200.h) instead of the
//    hop-sorted order; get_primal_variable_index() / get_bdd_index() describe it, and all
//    per-layer vectors of one solver share it (that is all callers rely on); reference_layer_order()
//    gives the permutation onto the reference's order for data that crosses between the two;
//  * node-level accessors of the reference's internal layout (get_lo_bdd_node_index, ...), the
//    sum-marginal functions of the learned solver are not provided (out of scope, SURVEY 2 rows
//    16-18); cereal save / load carry the library's own state blob (bddb200_save), not the
//    reference's member list; the protected members are gone, so the learned
//    subclass bdd_cuda_learned_mma does not build on top of this class.
#pragma once

#include <algorithm>
#include <array>
#include <cassert>
#include <climits>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include <thrust/device_ptr.h>
#include <thrust/device_vector.h>

#include "bdd_collection/bdd_collection.h"
#include "two_dimensional_variable_array.hxx"

#include "bdd_b200.h"

namespace LPMP {

    static constexpr int TOP_SINK_INDICATOR_CUDA = -1;
    static constexpr int BOT_SINK_INDICATOR_CUDA = -2;
    static constexpr int NUM_THREADS_CUDA = 256;

    namespace bddb200_detail {
        inline void check(int code)
        {
            if(code != BDDB200_OK)
                throw std::runtime_error(std::string("bdd_b200: ") + bddb200_last_error());
        }
        template<typename REAL> struct precision_of;
        template<> struct precision_of<float> { static constexpr int value = BDDB200_FLOAT; };
        template<> struct precision_of<double> { static constexpr int value = BDDB200_DOUBLE; };
    }

    template<typename REAL>
    class bdd_cuda_base {
        public:
            using value_type = REAL;
            using SOLVER_COSTS_VECS = std::tuple<thrust::device_vector<REAL>, thrust::device_vector<REAL>, thrust::device_vector<REAL>>;

            bdd_cuda_base() {}
            bdd_cuda_base(const BDD::bdd_collection& bdd_col) { construct(bdd_col, nullptr); }
            bdd_cuda_base(const BDD::bdd_collection& bdd_col, const std::vector<double>& costs_hi) { construct(bdd_col, &costs_hi); }
            ~bdd_cuda_base() { if(h_) bddb200_destroy(h_); }

            // construct_solver returns the solver by value into a std::variant (bdd_solver.cpp:130, 169-173)
            bdd_cuda_base(bdd_cuda_base&& o) noexcept { move_from(std::move(o)); }
            bdd_cuda_base& operator=(bdd_cuda_base&& o) noexcept { if(this != &o) { if(h_) bddb200_destroy(h_); h_ = nullptr; move_from(std::move(o)); } return *this; }
            bdd_cuda_base(const bdd_cuda_base& o) { copy_from(o); }
            bdd_cuda_base& operator=(const bdd_cuda_base& o) { if(this != &o) { if(h_) bddb200_destroy(h_); h_ = nullptr; copy_from(o); } return *this; }

            void flush_forward_states() { bddb200_flush_forward_states(h_); }
            void flush_backward_states() { bddb200_flush_backward_states(h_); }

            double lower_bound() { double lb; bddb200_detail::check(bddb200_lower_bound(h_, &lb)); return lb; }
            void lower_bound_per_bdd(thrust::device_ptr<REAL> lb_per_bdd) { bddb200_detail::check(bddb200_lower_bound_per_bdd(h_, thrust::raw_pointer_cast(lb_per_bdd))); sync(); }

            void update_costs(const std::vector<REAL>& cost_delta_0, const std::vector<REAL>& cost_delta_1)
            {
                bddb200_detail::check(bddb200_update_costs_host_real(h_, cost_delta_0.data(), cost_delta_0.size(), cost_delta_1.data(), cost_delta_1.size()));
            }
            void update_costs(const thrust::device_vector<REAL>& cost_delta_0, const thrust::device_vector<REAL>& cost_delta_1)
            {
                bddb200_detail::check(bddb200_update_costs_dev(h_, thrust::raw_pointer_cast(cost_delta_0.data()), cost_delta_0.size(),
                                                              thrust::raw_pointer_cast(cost_delta_1.data()), cost_delta_1.size()));
                sync();
            }
            template<typename REAL_arg>
            void update_costs(const thrust::device_ptr<const REAL_arg> cost_delta_0, const size_t delta_0_size,
                              const thrust::device_ptr<const REAL_arg> cost_delta_1, const size_t delta_1_size)
            {
                static_assert(std::is_same<REAL_arg, REAL>::value, "cost vectors must have the solver's REAL type");
                bddb200_detail::check(bddb200_update_costs_dev(h_, thrust::raw_pointer_cast(cost_delta_0), delta_0_size, thrust::raw_pointer_cast(cost_delta_1), delta_1_size));
                sync();
            }
            // host ranges of any arithmetic type, e.g. update_costs(c.begin(), c.begin(), c.begin(), c.end()) to add c to the hi costs: what
            // bdd_cuda<REAL> (src/bdd_cuda.cpp:63) and the reference's pybind constructor (bdd_cuda_parallel_mma_py.cu:42) call
            // (bdd_cuda_base.cu:476-503).  Variables behind the end of a range keep their costs.
            template<typename COST_ITERATOR>
            void update_costs(COST_ITERATOR cost_lo_begin, COST_ITERATOR cost_lo_end, COST_ITERATOR cost_hi_begin, COST_ITERATOR cost_hi_end)
            {
                const std::vector<double> lo(cost_lo_begin, cost_lo_end), hi(cost_hi_begin, cost_hi_end);
                bddb200_detail::check(bddb200_update_costs_host(h_, lo.data(), lo.size(), hi.data(), hi.size()));
            }
            void set_cost(const double c, const size_t var) { bddb200_detail::check(bddb200_set_cost(h_, c, var)); }

            // two_dim_variable_array<array<double,2>>: per variable the (mm_lo, mm_hi) of each of its BDDs (bdd_cuda_base.cu:753-786)
            two_dim_variable_array<std::array<double,2>> min_marginals()
            {
                auto [idx, lo, hi] = min_marginals_cuda(true);
                std::vector<REAL> h_lo(lo.size()), h_hi(hi.size());
                thrust::copy(lo.begin(), lo.end(), h_lo.begin());
                thrust::copy(hi.begin(), hi.end(), h_hi.begin());
                std::vector<int> h_idx(idx.size());
                thrust::copy(idx.begin(), idx.end(), h_idx.begin());
                two_dim_variable_array<std::array<double,2>> out;
                size_t pos = 0;
                const size_t n_inner = nr_layers() - nr_bdds();
                for(size_t var = 0; var < nr_variables(); ++var)
                {
                    std::vector<std::array<double,2>> mms;
                    while(pos < n_inner && (size_t)h_idx[pos] == var) { mms.push_back({(double)h_lo[pos], (double)h_hi[pos]}); ++pos; }
                    out.push_back(mms.begin(), mms.end());
                }
                return out;
            }
            std::tuple<thrust::device_vector<int>, thrust::device_vector<REAL>, thrust::device_vector<REAL>> min_marginals_cuda(bool get_sorted = true)
            {
                thrust::device_vector<int> idx(nr_layers());
                thrust::device_vector<REAL> lo(nr_layers()), hi(nr_layers());
                bddb200_detail::check(bddb200_min_marginals(h_, get_sorted ? 1 : 0, thrust::raw_pointer_cast(idx.data()),
                                                           thrust::raw_pointer_cast(lo.data()), thrust::raw_pointer_cast(hi.data())));
                sync();
                return {std::move(idx), std::move(lo), std::move(hi)};
            }

            thrust::device_vector<char> bdds_solution_vec()
            {
                thrust::device_vector<char> sol(nr_layers());
                bddb200_detail::check(bddb200_bdds_solution(h_, thrust::raw_pointer_cast(sol.data())));
                sync();
                return sol;
            }
            two_dim_variable_array<REAL> bdds_solution()
            {
                const thrust::device_vector<char> sol = bdds_solution_vec();
                std::vector<char> h_sol(sol.size());
                thrust::copy(sol.begin(), sol.end(), h_sol.begin());
                std::vector<std::vector<REAL>> per_var(nr_variables());
                for(size_t l = 0; l < h_sol.size(); ++l)
                    if(primal_variable_index_host_[l] != INT_MAX) per_var[primal_variable_index_host_[l]].push_back((REAL)h_sol[l]);
                two_dim_variable_array<REAL> out;
                for(const auto& v : per_var) out.push_back(v.begin(), v.end());
                return out;
            }

            std::vector<REAL> get_primal_objective_vector_host()
            {
                std::vector<double> obj(nr_variables());
                bddb200_detail::check(bddb200_primal_objective_host(h_, obj.data()));
                return std::vector<REAL>(obj.begin(), obj.end());
            }

            size_t nr_variables() const { return nr_vars_; }
            size_t nr_variables(const size_t bdd_nr) const { assert(bdd_nr < nr_bdds()); return nr_vars_; }
            size_t nr_bdds() const { return nr_bdds_; }
            size_t nr_bdds(const size_t var) const { assert(var < nr_variables()); return num_bdds_per_var_host_[var]; }
            size_t nr_layers() const { return cum_nr_layers_per_hop_dist_.empty() ? 0 : cum_nr_layers_per_hop_dist_.back(); }
            size_t nr_layers(const int hop_index) const { return cum_nr_layers_per_hop_dist_[hop_index] - (hop_index > 0 ? cum_nr_layers_per_hop_dist_[hop_index - 1] : 0); }
            size_t nr_bdd_nodes() const { return cum_nr_bdd_nodes_per_hop_dist_.empty() ? 0 : cum_nr_bdd_nodes_per_hop_dist_.back(); }
            size_t nr_bdd_nodes(const int hop_index) const { return cum_nr_bdd_nodes_per_hop_dist_[hop_index] - (hop_index > 0 ? cum_nr_bdd_nodes_per_hop_dist_[hop_index - 1] : 0); }
            size_t nr_hops() const { return cum_nr_layers_per_hop_dist_.size() - 1; }   // ignores terminal nodes

            void forward_run() { bddb200_detail::check(bddb200_forward_run(h_)); }
            // the reference returns per-node path costs when asked; nothing outside the class consumes them
            std::tuple<thrust::device_vector<REAL>, thrust::device_vector<REAL>> backward_run(bool compute_path_costs = true)
            {
                (void)compute_path_costs;
                bddb200_detail::check(bddb200_backward_run(h_));
                return {};
            }

            std::tuple<thrust::device_vector<int>, thrust::device_vector<int>> var_constraint_indices() const
            {
                return {get_primal_variable_index(), get_bdd_index()};
            }

            void get_solver_costs(thrust::device_ptr<REAL> lo_cost_out_ptr, thrust::device_ptr<REAL> hi_cost_out_ptr, thrust::device_ptr<REAL> deferred_mm_diff_out_ptr) const
            {
                bddb200_detail::check(bddb200_get_solver_costs(h_, thrust::raw_pointer_cast(lo_cost_out_ptr), thrust::raw_pointer_cast(hi_cost_out_ptr),
                                                              thrust::raw_pointer_cast(deferred_mm_diff_out_ptr)));
                sync();
            }
            SOLVER_COSTS_VECS get_solver_costs() const
            {
                thrust::device_vector<REAL> lo(nr_layers()), hi(nr_layers()), mm(nr_layers());
                get_solver_costs(lo.data(), hi.data(), mm.data());
                return {std::move(lo), std::move(hi), std::move(mm)};
            }
            void set_solver_costs(const thrust::device_ptr<const REAL> lo_costs, const thrust::device_ptr<const REAL> hi_costs, const thrust::device_ptr<const REAL> deferred_mm_diff)
            {
                bddb200_detail::check(bddb200_set_solver_costs(h_, thrust::raw_pointer_cast(lo_costs), thrust::raw_pointer_cast(hi_costs), thrust::raw_pointer_cast(deferred_mm_diff)));
                sync();
            }
            void set_solver_costs(const SOLVER_COSTS_VECS& costs)
            {
                set_solver_costs(std::get<0>(costs).data(), std::get<1>(costs).data(), std::get<2>(costs).data());
            }

            const thrust::device_vector<int> get_primal_variable_index() const { return thrust::device_vector<int>(primal_variable_index_host_.begin(), primal_variable_index_host_.end()); }
            const thrust::device_vector<int> get_bdd_index() const { return thrust::device_vector<int>(bdd_index_host_.begin(), bdd_index_host_.end()); }
            const thrust::device_vector<int>& get_num_bdds_per_var() const { return num_bdds_per_var_; }

            // perm[k] = position, in this class's BDD-major layer order, of the layer the reference class keeps at position k.
            // The reference sorts its nodes by (hop distance, primal variable, BDD) and compresses equal keys to layers
            // (bdd_cuda_base.cu:146-188, :240-285; both sinks of a BDD share the hop after its last variable and the variable INT_MAX,
            // :113-129).  thrust::gather(perm, v) turns a per-layer vector of this class into the reference's order, thrust::scatter
            // brings one saved from the reference back (pinned against the reference's own CUDA build in tests/test_layer_order_gpu.py).
            std::vector<int> reference_layer_order() const
            {
                const size_t n = bdd_index_host_.size();
                std::vector<int> hop(n), perm(n);
                for(size_t k = 0, first = 0; k < n; ++k)
                {
                    if(k > 0 && bdd_index_host_[k] != bdd_index_host_[k - 1]) first = k;
                    hop[k] = int(k - first);
                    perm[k] = int(k);
                }
                std::sort(perm.begin(), perm.end(), [&](int a, int b) {
                    return std::make_tuple(hop[a], primal_variable_index_host_[a], bdd_index_host_[a]) < std::make_tuple(hop[b], primal_variable_index_host_[b], bdd_index_host_[b]);
                });
                return perm;
            }

            void distribute_delta() { bddb200_detail::check(bddb200_distribute_delta(h_)); }

            void make_dual_feasible(thrust::device_vector<REAL>& d) const
            {
                assert(d.size() == nr_layers());
                bddb200_detail::check(bddb200_make_dual_feasible(h_, thrust::raw_pointer_cast(d.data())));
                sync();
            }

            const std::vector<int>& get_cum_nr_bdd_nodes_per_hop_dist() const { return cum_nr_bdd_nodes_per_hop_dist_; }
            const std::vector<int>& get_cum_nr_layers_per_hop_dist() const { return cum_nr_layers_per_hop_dist_; }

            void terminal_layer_indices(thrust::device_ptr<int> indices) const
            {
                std::vector<int> t;
                for(size_t l = 0; l < primal_variable_index_host_.size(); ++l)
                    if(primal_variable_index_host_[l] == INT_MAX) t.push_back((int)l);
                thrust::copy(t.begin(), t.end(), indices);
            }

            bddb200_solver* native_handle() const { return h_; }

        protected:
            void sync() const { bddb200_detail::check(bddb200_synchronize(h_)); }   // the reference's thrust calls are synchronous

            bddb200_solver* h_ = nullptr;
            size_t nr_vars_ = 0, nr_bdds_ = 0;
            std::vector<int> num_bdds_per_var_host_, primal_variable_index_host_, bdd_index_host_;
            thrust::device_vector<int> num_bdds_per_var_;
            std::vector<int> cum_nr_bdd_nodes_per_hop_dist_, cum_nr_layers_per_hop_dist_;

        private:
            void construct(const BDD::bdd_collection& bdd_col, const std::vector<double>* costs_hi)
            {
                static_assert(sizeof(BDD::bdd_instruction) == sizeof(bddb200_instruction), "bdd_instruction layout");
                const size_t nb = bdd_col.nr_bdds();
                if(nb == 0) throw std::runtime_error("bdd_b200: empty bdd collection");
                std::vector<size_t> delims(nb + 1);
                for(size_t b = 0; b < nb; ++b) delims[b] = bdd_col.offset(b);
                delims[nb] = delims[nb - 1] + bdd_col.nr_bdd_nodes(nb - 1);
                const BDD::bdd_instruction* first = &*bdd_col.get_bdd_instructions(0).first - delims[0];
                // hop statistics (bdd_cuda_base.cu:86-144, 240-307): layer k of every BDD is hop k
                std::vector<int> layers_per_hop, nodes_per_hop;
                for(size_t b = 0; b < nb; ++b)
                {
                    size_t hop = 0, prev = first[delims[b]].index;
                    for(size_t i = delims[b]; i < delims[b + 1]; ++i)
                    {
                        const bool terminal = first[i].index >= (size_t)-2;
                        if(!terminal && first[i].index != prev) { ++hop; prev = first[i].index; }
                        const size_t at = terminal ? hop + 1 : hop;
                        if(nodes_per_hop.size() <= at) { nodes_per_hop.resize(at + 1, 0); layers_per_hop.resize(at + 1, 0); }
                        nodes_per_hop[at]++;
                        if(!terminal && (i == delims[b] || first[i - 1].index != first[i].index)) layers_per_hop[at]++;
                    }
                    layers_per_hop[hop + 1]++;    // the BDD's terminal layer
                }
                cum_nr_bdd_nodes_per_hop_dist_.assign(nodes_per_hop.size(), 0);
                cum_nr_layers_per_hop_dist_.assign(nodes_per_hop.size(), 0);
                int cn = 0, cl = 0;
                for(size_t k = 0; k < nodes_per_hop.size(); ++k) { cn += nodes_per_hop[k]; cl += layers_per_hop[k]; cum_nr_bdd_nodes_per_hop_dist_[k] = cn; cum_nr_layers_per_hop_dist_[k] = cl; }

                if(std::getenv("BDDB200_DEBUG")) std::fprintf(stderr, "[shim] create nb=%zu n=%zu hops=%zu\n", nb, delims[nb], nodes_per_hop.size());
                bddb200_detail::check(bddb200_create(reinterpret_cast<const bddb200_instruction*>(first), delims[nb], delims.data(), nb,
                                                    costs_hi ? costs_hi->data() : nullptr, costs_hi ? costs_hi->size() : 0,
                                                    bddb200_detail::precision_of<REAL>::value, nullptr, &h_));
                if(std::getenv("BDDB200_DEBUG")) std::fprintf(stderr, "[shim] created\n");
                fetch_sizes();
                if(std::getenv("BDDB200_DEBUG")) std::fprintf(stderr, "[shim] sizes fetched\n");
            }
        public:
            // cereal interface of the reference class (include/bdd_solver/bdd_cuda_base.h:166-170, bdd_cuda_base.cu:1486-1544): the archive
            // receives the two hop tables and the library's state blob (bddb200_save); load() rebuilds the solver from it on device 0
            template<class Archive>
            void save(Archive& archive) const
            {
                std::vector<char> blob;
                if(h_)
                {
                    size_t n = 0;
                    bddb200_detail::check(bddb200_save_size(h_, &n));
                    blob.resize(n);
                    bddb200_detail::check(bddb200_save(h_, blob.data(), blob.size(), &n));
                }
                archive(cum_nr_bdd_nodes_per_hop_dist_, cum_nr_layers_per_hop_dist_, blob);
            }
            template<class Archive>
            void load(Archive& archive)
            {
                std::vector<char> blob;
                archive(cum_nr_bdd_nodes_per_hop_dist_, cum_nr_layers_per_hop_dist_, blob);
                if(h_) { bddb200_destroy(h_); h_ = nullptr; }
                if(blob.empty()) return;
                bddb200_detail::check(bddb200_load(blob.data(), blob.size(), 0, &h_));
                if(bddb200_precision_of(h_) != bddb200_detail::precision_of<REAL>::value)
                {
                    bddb200_destroy(h_); h_ = nullptr;
                    throw std::runtime_error("bdd_b200: archive holds a solver of the other precision");
                }
                fetch_sizes();
            }
        protected:
            void fetch_sizes()
            {
                nr_vars_ = bddb200_nr_variables(h_); nr_bdds_ = bddb200_nr_bdds(h_);
                num_bdds_per_var_host_.resize(nr_vars_);
                bddb200_detail::check(bddb200_nr_bdds_per_var(h_, num_bdds_per_var_host_.data()));
                num_bdds_per_var_ = thrust::device_vector<int>(num_bdds_per_var_host_.begin(), num_bdds_per_var_host_.end());
                const size_t nl = bddb200_nr_layers(h_);
                primal_variable_index_host_.resize(nl); bdd_index_host_.resize(nl);
                bddb200_detail::check(bddb200_layer_primal_indices(h_, primal_variable_index_host_.data()));
                bddb200_detail::check(bddb200_layer_bdd_indices(h_, bdd_index_host_.data()));
            }
            void move_from(bdd_cuda_base&& o)
            {
                h_ = o.h_; o.h_ = nullptr;
                nr_vars_ = o.nr_vars_; nr_bdds_ = o.nr_bdds_;
                num_bdds_per_var_host_ = std::move(o.num_bdds_per_var_host_);
                primal_variable_index_host_ = std::move(o.primal_variable_index_host_);
                bdd_index_host_ = std::move(o.bdd_index_host_);
                num_bdds_per_var_ = std::move(o.num_bdds_per_var_);
                cum_nr_bdd_nodes_per_hop_dist_ = std::move(o.cum_nr_bdd_nodes_per_hop_dist_);
                cum_nr_layers_per_hop_dist_ = std::move(o.cum_nr_layers_per_hop_dist_);
            }
            void copy_from(const bdd_cuda_base& o)
            {
                if(o.h_) bddb200_detail::check(bddb200_clone(o.h_, &h_));
                nr_vars_ = o.nr_vars_; nr_bdds_ = o.nr_bdds_;
                num_bdds_per_var_host_ = o.num_bdds_per_var_host_;
                primal_variable_index_host_ = o.primal_variable_index_host_;
                bdd_index_host_ = o.bdd_index_host_;
                num_bdds_per_var_ = o.num_bdds_per_var_;
                cum_nr_bdd_nodes_per_hop_dist_ = o.cum_nr_bdd_nodes_per_hop_dist_;
                cum_nr_layers_per_hop_dist_ = o.cum_nr_layers_per_hop_dist_;
            }
    };

    template<typename REAL>
    class bdd_cuda_parallel_mma : public bdd_cuda_base<REAL> {
        public:
            void init() {}
            bdd_cuda_parallel_mma() {}
            bdd_cuda_parallel_mma(const BDD::bdd_collection& bdd_col) : bdd_cuda_base<REAL>(bdd_col) {}
            bdd_cuda_parallel_mma(const BDD::bdd_collection& bdd_col, const std::vector<double>& costs) : bdd_cuda_base<REAL>(bdd_col, costs) {}

            void iteration(const REAL omega = 0.5) { bddb200_detail::check(bddb200_iteration(this->h_, omega)); }

            void forward_mm(const REAL omega, thrust::device_vector<REAL>& delta_lo_hi)
            {
                assert(delta_lo_hi.size() == 2 * this->nr_variables());
                bddb200_detail::check(bddb200_forward_mm(this->h_, omega, thrust::raw_pointer_cast(delta_lo_hi.data())));
                this->sync();
            }
            void backward_mm(const REAL omega, thrust::device_vector<REAL>& delta_lo_hi)
            {
                assert(delta_lo_hi.size() == 2 * this->nr_variables());
                bddb200_detail::check(bddb200_backward_mm(this->h_, omega, thrust::raw_pointer_cast(delta_lo_hi.data())));
                this->sync();
            }
            void normalize_delta(thrust::device_vector<REAL>& delta_lo_hi) const
            {
                bddb200_detail::check(bddb200_normalize_delta(this->h_, thrust::raw_pointer_cast(delta_lo_hi.data())));
                this->sync();
            }
            thrust::device_vector<REAL> net_solver_costs() const
            {
                thrust::device_vector<REAL> out(this->nr_layers());
                bddb200_detail::check(bddb200_net_solver_costs(this->h_, thrust::raw_pointer_cast(out.data())));
                this->sync();
                return out;
            }
            void gradient_step(const thrust::device_vector<REAL>& g, double step_size)
            {
                assert(g.size() == this->nr_layers());
                bddb200_detail::check(bddb200_gradient_step(this->h_, thrust::raw_pointer_cast(g.data()), step_size));
                this->sync();      // `g` is caller-owned: the reference call is synchronous, the caller may overwrite it on return
            }
    };

}
