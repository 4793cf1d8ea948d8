// tests/cpp/test_dropin.cu -- the reference's own CPU-vs-GPU equivalence protocol
// (test/test_cuda_parallel_mma.cu:13-103) and known-answer test
// (test/test_bdd_cuda_parallel_mma.cu:197-247) compiled against the DROP-IN class
//   #include "bdd_solver/bdd_cuda_parallel_mma.h"   -> bdd_b200/csrc/host/bdd_solver/...
// with everything else taken from the reference tree where it lies: BDD::bdd_collection, the
// inequality -> BDD converter, the CPU solver bdd_parallel_mma_base<...> and the driver loop
// run_solver (include/run_solver_util.h).  TEST INFRASTRUCTURE: built by `make -C oracle dropin`
// where /root/reference exists (output oracle/_ref/test_dropin), run by tests/test_dropin_gpu.py.
//
// The reference tests parse LP strings with PEGTL (absent here); the same problems are given as
// constraint lists and converted by the reference's own converter.
#include "bdd_solver/bdd_parallel_mma_base.h"
#include "bdd_solver/bdd_cuda_parallel_mma.h"
#include "bdd_solver/lbfgs.h"                      // -> bdd_b200/csrc/host/bdd_solver/lbfgs.h (+ the reference header via #include_next)
#include "bdd_solver/bdd_branch_instruction.h"
#include "bdd_conversion/convert_pb_to_bdd.h"
#include "bdd_manager/bdd_mgr.h"
#include "run_solver_util.h"
#include <cereal/archives/binary.hpp>
#include <cmath>
#include <cstdio>
#include <iostream>
#include <sstream>
#include <numeric>
#include <random>
#include <variant>

using namespace LPMP;

static int failures = 0;
#define STEP(msg) do { if(std::getenv("BDDB200_DEBUG")) std::fprintf(stderr, "[step] %s\n", msg); } while(0)
#define CHECK(cond) do { if(!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } } while(0)

struct Constraint { std::vector<int> coeffs; std::vector<size_t> vars; int ineq; int rhs; };   // ineq: 0 <=, 1 >=, 2 =
struct Problem { const char* name; size_t nr_vars; std::vector<double> objective; std::vector<Constraint> constraints; double expected_lb; };

// bdd_preprocessor.cpp:175-228: simplex shortcut, else convert -> add_bdd -> reorder -> make_qbdd -> rebase
static BDD::bdd_collection build_collection(const Problem& p)
{
    BDD::bdd_collection col;
    BDD::bdd_mgr mgr;
    bdd_converter conv(mgr);
    for(const Constraint& c : p.constraints)
    {
        bool simplex = c.ineq == 2 && c.rhs != 0;
        for(int a : c.coeffs) if(a != c.rhs) simplex = false;
        if(simplex)
        {
            const size_t bdd_nr = col.simplex_constraint(c.vars.size());
            col.rebase(bdd_nr, c.vars.begin(), c.vars.end());
            continue;
        }
        const ILP_input::inequality_type it = c.ineq == 0 ? ILP_input::inequality_type::smaller_equal
            : (c.ineq == 1 ? ILP_input::inequality_type::greater_equal : ILP_input::inequality_type::equal);
        BDD::node_ref bdd = conv.convert_to_bdd(c.coeffs, it, c.rhs);
        size_t bdd_nr = col.add_bdd(bdd);
        col.reorder(bdd_nr);
        if(!col.is_qbdd(bdd_nr)) { col.make_qbdd(bdd_nr); col.remove(bdd_nr); }
        col.rebase(bdd_nr, c.vars.begin(), c.vars.end());
    }
    return col;
}

static Problem matching_3x3()
{
    Problem p{"matching_3x3", 9, {-2, -1, -1, -1, -2, -1, -1, -1, -2}, {}, -6.0};
    for(size_t r = 0; r < 3; ++r) p.constraints.push_back({{1, 1, 1}, {3 * r, 3 * r + 1, 3 * r + 2}, 2, 1});
    for(size_t c = 0; c < 3; ++c) p.constraints.push_back({{1, 1, 1}, {c, c + 3, c + 6}, 2, 1});
    return p;
}
// test/test_bdd_cuda_parallel_mma.cu:23-38 (variables numbered by first appearance, objective first)
static Problem short_chain_shuffled()
{   // mu_2_1=0 mu_10=1 mu_1_1=2 mu_11=3 mu_1_0=4 mu_00=5 mu_01=6 mu_2_0=7
    Problem p{"short_chain_shuffled", 8, {1, 1, 0, 0, -1, 1, 2, 2}, {}, 1.0};
    p.constraints.push_back({{1, 1}, {4, 2}, 2, 1});
    p.constraints.push_back({{1, 1}, {7, 0}, 2, 1});
    p.constraints.push_back({{1, 1, 1, 1}, {5, 1, 6, 3}, 2, 1});
    p.constraints.push_back({{1, -1, -1}, {4, 5, 6}, 2, 0});
    p.constraints.push_back({{1, -1, -1}, {2, 1, 3}, 2, 0});
    p.constraints.push_back({{1, -1, -1}, {7, 5, 1}, 2, 0});
    p.constraints.push_back({{1, -1, -1}, {0, 6, 3}, 2, 0});
    return p;
}
// a covering / knapsack mix with wider BDDs
static Problem knapsack_mix()
{
    Problem p{"knapsack_mix", 10, {3, -2, 4, 1, -5, 2, 2, -1, 6, -3}, {}, std::nan("")};
    p.constraints.push_back({{2, 3, 4, 5, 6}, {0, 1, 2, 3, 4}, 0, 9});
    p.constraints.push_back({{1, 1, 1, 1}, {4, 5, 6, 7}, 1, 2});
    p.constraints.push_back({{3, -2, 5, 1, 2, 4}, {2, 5, 6, 7, 8, 9}, 0, 6});
    p.constraints.push_back({{1, 1, 1}, {0, 8, 9}, 2, 1});
    p.constraints.push_back({{1, 2, 3, 1}, {1, 3, 7, 9}, 1, 2});
    return p;
}

// test/test_cuda_parallel_mma.cu:13-103
static void test_equivalence(const Problem& p, const bool with_additional_gaps)
{
    BDD::bdd_collection bdd_col = build_collection(p);
    std::mt19937 gen(0);
    std::uniform_int_distribution<> distrib(0, 6);
    std::vector<size_t> var_map;
    if(with_additional_gaps)
    {
        var_map.push_back(distrib(gen));
        for(size_t i = 1; i < p.nr_vars; ++i) var_map.push_back(var_map.back() + 1 + distrib(gen));
    }
    else { var_map.resize(p.nr_vars); std::iota(var_map.begin(), var_map.end(), 0); }
    bdd_col.rebase(var_map.begin(), var_map.end());
    std::vector<double> mapped_obj(var_map.back() + 1, 0.0);
    for(size_t i = 0; i < p.nr_vars; ++i) mapped_obj[var_map[i]] = p.objective[i];

    // before any cost: both solvers built from the collection alone
    {
        bdd_parallel_mma_base<bdd_branch_instruction<double, uint16_t>> cpu0(bdd_col);
        bdd_cuda_parallel_mma<double> gpu0(bdd_col);
        CHECK(std::abs(cpu0.lower_bound() - gpu0.lower_bound()) < 1e-6);
        // the GPU class takes costs after construction too (test_cuda_parallel_mma.cu:60)
        gpu0.update_costs({}, mapped_obj);
        bdd_cuda_parallel_mma<double> gpu1(bdd_col, mapped_obj);
        CHECK(std::abs(gpu0.lower_bound() - gpu1.lower_bound()) < 1e-12);
    }
    // The reference test calls parallel_mma.update_costs({}, mapped_obj) here; at this commit that
    // overload passes its iterators in the wrong order (bdd_parallel_mma_base.cpp:623) and crashes,
    // so the CPU solver gets its costs through the two-argument constructor (:30) instead.
    bdd_parallel_mma_base<bdd_branch_instruction<double, uint16_t>> parallel_mma(bdd_col, mapped_obj);
    bdd_cuda_parallel_mma<double> cuda_mma(bdd_col, mapped_obj);
    STEP("constructed");

    CHECK(parallel_mma.nr_variables() == var_map.back() + 1);
    CHECK(parallel_mma.nr_variables() == cuda_mma.nr_variables());
    CHECK(parallel_mma.nr_bdds() == cuda_mma.nr_bdds());
    for(size_t i = 0; i < parallel_mma.nr_variables(); ++i) CHECK(parallel_mma.nr_bdds(i) == cuda_mma.nr_bdds(i));
    CHECK(std::abs(parallel_mma.lower_bound() - cuda_mma.lower_bound()) < 1e-6);

    std::vector<std::array<double, 2>> cpu_delta(parallel_mma.nr_variables(), std::array<double, 2>{0.0, 0.0});
    thrust::device_vector<double> cuda_delta(2 * parallel_mma.nr_variables(), 0.0);
    for(size_t iter = 0; iter < 10; ++iter)
    {
        parallel_mma.forward_mm(0.5, cpu_delta);
        cuda_mma.forward_mm(0.5, cuda_delta);
        for(size_t i = 0; i < parallel_mma.nr_variables(); ++i)
        {
            CHECK(std::abs(cpu_delta[i][0] - cuda_delta[2 * i]) < 1e-6);
            CHECK(std::abs(cpu_delta[i][1] - cuda_delta[2 * i + 1]) < 1e-6);
        }
        parallel_mma.backward_mm(0.5, cpu_delta);
        cuda_mma.backward_mm(0.5, cuda_delta);
        for(size_t i = 0; i < parallel_mma.nr_variables(); ++i)
        {
            CHECK(std::abs(cpu_delta[i][0] - cuda_delta[2 * i]) < 1e-6);
            CHECK(std::abs(cpu_delta[i][1] - cuda_delta[2 * i + 1]) < 1e-6);
        }
        CHECK(std::abs(parallel_mma.lower_bound() - cuda_mma.lower_bound()) < 1e-6);
    }
    std::printf("equivalence %-22s gaps=%d: lb cpu %.9f gpu %.9f\n", p.name, (int)with_additional_gaps, parallel_mma.lower_bound(), cuda_mma.lower_bound());
}

// test/test_bdd_cuda_parallel_mma.cu:197-247
template<typename REAL>
static void test_known_answer(const Problem& p, const double tol)
{
    BDD::bdd_collection bdd_col = build_collection(p);
    bdd_cuda_parallel_mma<REAL> solver(bdd_col);
    for(size_t i = 0; i < p.nr_vars; ++i) solver.set_cost(p.objective[i], i);
    {
        const std::vector<REAL> obj = solver.get_primal_objective_vector_host();
        for(size_t i = 0; i < p.nr_vars; ++i) CHECK(std::abs(obj[i] - p.objective[i]) <= tol);
    }
    for(size_t iter = 0; iter < 200; ++iter) solver.iteration();
    solver.distribute_delta();
    CHECK(std::abs(solver.lower_bound() - p.expected_lb) <= tol);
    {
        const std::vector<REAL> obj = solver.get_primal_objective_vector_host();
        for(size_t i = 0; i < p.nr_vars; ++i) CHECK(std::abs(obj[i] - p.objective[i]) <= std::max(tol, 1e-12) * 10);
    }
    // min_marginals(): per variable one pair per BDD, mm_lo / mm_hi >= lb of that BDD
    const auto mms = solver.min_marginals();
    CHECK(mms.size() == solver.nr_variables());
    for(size_t v = 0; v < mms.size(); ++v) CHECK(mms.size(v) == solver.nr_bdds(v));
    std::printf("known answer %-22s %s: lb %.12f (expected %.1f)\n", p.name, sizeof(REAL) == 8 ? "double" : "float", solver.lower_bound(), p.expected_lb);
}

// the reference's driver loop and its by-value std::variant handling (bdd_solver.h:58-69, bdd_solver.cpp:130, 303-305)
static void test_run_solver_and_variant(const Problem& p)
{
    BDD::bdd_collection bdd_col = build_collection(p);
    using variant_t = std::variant<bdd_parallel_mma_base<bdd_branch_instruction<double, uint16_t>>, bdd_cuda_parallel_mma<float>, bdd_cuda_parallel_mma<double>>;
    auto construct = [&]() -> variant_t { return bdd_cuda_parallel_mma<double>(bdd_col, p.objective); };
    variant_t solver = construct();
    std::visit([&](auto&& s) { run_solver(s, 1000, 1e-9, 1e-12, 3600.0); }, solver);
    const double lb = std::visit([&](auto&& s) { return s.lower_bound(); }, solver);
    CHECK(std::abs(lb - p.expected_lb) < 1e-3);
    // copies are deep: iterating the copy leaves the original untouched
    bdd_cuda_parallel_mma<double> a(bdd_col, p.objective);
    a.iteration();
    bdd_cuda_parallel_mma<double> b(a);
    const double lb_a = a.lower_bound();
    for(int i = 0; i < 5; ++i) b.iteration();
    CHECK(a.lower_bound() == lb_a);
    CHECK(b.lower_bound() >= lb_a - 1e-12);
    for(int i = 0; i < 5; ++i) a.iteration();
    CHECK(a.lower_bound() == b.lower_bound());
    std::printf("run_solver / variant / copy %-14s: lb %.9f\n", p.name, lb);
}

// cereal round trip as the reference's pybind module pickles a solver (src/bdd_solver/bdd_cuda_parallel_mma_py.cu:15-38): archive(solver)
// calls the member save / load templates
static void test_save_load(const Problem& p)
{
    BDD::bdd_collection bdd_col = build_collection(p);
    bdd_cuda_parallel_mma<double> a(bdd_col, p.objective);
    for(int i = 0; i < 3; ++i) a.iteration();
    std::stringstream ss;
    {
        cereal::BinaryOutputArchive out(ss);
        a.save(out);
    }
    bdd_cuda_parallel_mma<double> b;
    {
        cereal::BinaryInputArchive in(ss);
        b.load(in);
        b.init();
    }
    CHECK(b.nr_variables() == a.nr_variables() && b.nr_bdds() == a.nr_bdds() && b.nr_layers() == a.nr_layers() && b.nr_hops() == a.nr_hops());
    CHECK(b.lower_bound() == a.lower_bound());
    for(int i = 0; i < 200; ++i) { a.iteration(); b.iteration(); }
    CHECK(std::abs(a.lower_bound() - b.lower_bound()) < 1e-9);
    CHECK(std::abs(b.lower_bound() - p.expected_lb) < 1e-3);
    std::printf("cereal save / load        %-14s: lb %.9f\n", p.name, b.lower_bound());
}

// L-BFGS support surface as lbfgs<> uses it (include/bdd_solver/lbfgs.h:22-27, src/bdd_solver/lbfgs_impl.h)
static void test_lbfgs_surface(const Problem& p)
{
    BDD::bdd_collection bdd_col = build_collection(p);
    bdd_cuda_parallel_mma<double> s(bdd_col, p.objective);
    for(int i = 0; i < 3; ++i) s.iteration();
    thrust::device_vector<char> sol = s.bdds_solution_vec();
    thrust::device_vector<double> net = s.net_solver_costs();
    CHECK(sol.size() == s.nr_layers() && net.size() == s.nr_layers());
    thrust::device_vector<double> d(s.nr_layers(), 1.0);
    s.make_dual_feasible(d);                       // constant per variable -> all zero after mean removal
    std::vector<double> hd(d.size());
    thrust::copy(d.begin(), d.end(), hd.begin());
    for(double x : hd) CHECK(std::abs(x) < 1e-12);
    const double lb0 = s.lower_bound();
    s.gradient_step(d, 0.1);                        // zero direction: bound unchanged
    CHECK(std::abs(s.lower_bound() - lb0) < 1e-12);
    auto [idx, lo, hi] = s.min_marginals_cuda(true);
    CHECK(idx.size() == s.nr_layers());
    std::printf("lbfgs surface %-22s ok\n", p.name);
}

// The reference's GPU L-BFGS type (bdd_solver.h:60-61) resolves to the device implementation; it runs inside the reference's run_solver
// and std::variant like any other solver, keeps the bound monotone and ends at or above plain MMA's bound.
static void test_lbfgs_wrapper(const Problem& p)
{
    using lbfgs_type = lbfgs<bdd_cuda_parallel_mma<double>, thrust::device_vector<double>, double, thrust::device_vector<char>, true>;
    BDD::bdd_collection bdd_col = build_collection(p);
    std::variant<bdd_cuda_parallel_mma<double>, lbfgs_type> v = lbfgs_type(bdd_col, p.objective, 5, 1e-3, 1e-6, 0.8, 1.1);
    bdd_cuda_parallel_mma<double> plain(bdd_col, p.objective);
    double prev = std::visit([](auto& s) { return s.lower_bound(); }, v);
    for(int i = 0; i < 25; ++i)
    {
        std::visit([](auto& s) { s.iteration(); }, v);
        plain.iteration();
        const double lb = std::visit([](auto& s) { return s.lower_bound(); }, v);
        CHECK(lb >= prev - 1e-9);
        prev = lb;
    }
    CHECK(prev >= plain.lower_bound() - 1e-6);
    lbfgs_type copy = std::get<lbfgs_type>(v);                       // copy construction: own solver state, empty history
    copy.iteration();
    CHECK(copy.lower_bound() >= prev - 1e-9);
    std::visit([&](auto& s) { run_solver(s, 50, 1e-9, 1e-12, 3600.0); }, v);
    CHECK(std::abs(std::visit([](auto& s) { return s.lower_bound(); }, v) - p.expected_lb) < 1e-6);
    std::printf("lbfgs wrapper %-22s ok\n", p.name);
}

int main()
{
    std::setvbuf(stdout, nullptr, _IONBF, 0);
    const Problem problems[] = {matching_3x3(), short_chain_shuffled(), knapsack_mix()};
    for(const Problem& p : problems)
    {
        test_equivalence(p, false);
        test_equivalence(p, true);
    }
    test_known_answer<double>(problems[0], 1e-12);
    test_known_answer<double>(problems[1], 1e-12);
    test_known_answer<float>(problems[0], 1e-4);
    test_run_solver_and_variant(problems[1]);
    test_save_load(problems[1]);
    test_save_load(problems[0]);
    test_lbfgs_surface(problems[1]);
    test_lbfgs_wrapper(problems[1]);
    test_lbfgs_wrapper(problems[0]);
    std::printf(failures == 0 ? "ALL OK\n" : "%d FAILURES\n", failures);
    return failures == 0 ? 0 : 1;
}
