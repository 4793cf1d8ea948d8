// bdd_b200/csrc/host/bdd_solver_native.hpp -- C++ host side of the JSON-config driver for the GPU solvers.
//
// Mirrors LPMP::bdd_solver (include/bdd_solver/bdd_solver.h:45-103, src/bdd_solver/bdd_solver.cpp) for the relaxation solvers this build
// provides -- "cuda parallel mma" and the three spellings of its L-BFGS wrapper -- over the C ABI of libbdd_b200 (include/bdd_b200.h):
//   read_config (:468-475)  read_ILP (:44-66)  transform_to_BDDs (:112-123)  construct_solver (:130-267)  solve_dual (:277-309, the
//   run_solver loop of include/run_solver_util.h:10-77 with its per-iteration log line)  perturbation_rounding (:318-380)  solve (:477-495).
// The reference reads .lp files with a PEGTL grammar (src/ILP/ILP_parser.cpp:25-160) and turns inequalities into BDDs through its BDD
// manager (src/bdd_conversion, src/bdd_manager); neither library is part of this build: host/lp_reader.hpp is a small reader for the
// same LP subset, host/bdd_collection.hpp builds the quasi-reduced BDD of a linear 0/1 constraint directly (dynamic programme over partial
// sums, states with equal sub-functions merged bottom-up: the canonical form the reference reaches through lineq_bdd -> bdd_mgr -> reorder
// -> make_qbdd, src/bdd_conversion/bdd_preprocessor.cpp:199-215).  Same code as bdd_b200/lp.py and bdd_b200/instances.py:qbdd_template, in C++.
// Deviations from the reference, as in the Python driver: "precision" means what it says (the reference swaps float and
// double, bdd_solver.cpp:167-174), the README spelling "lbfgs cuda parallel mma" is accepted, GPU rounding works for all four solver types.
#pragma once

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include <nlohmann/json.hpp>

#include "../../../include/bdd_b200.h"
#include "bdd_collection.hpp"
#include "lp_reader.hpp"

namespace bddb200_host {

// ---------------------------------------------------------------------------------------------------------------- driver
class bdd_solver {
public:
    using json = nlohmann::json;

    bdd_solver() {}
    explicit bdd_solver(const std::string& config) { config_ = read_config(config); }
    explicit bdd_solver(const json& config) : config_(config) {}
    bdd_solver(const bdd_solver&) = delete;
    bdd_solver& operator=(const bdd_solver&) = delete;
    ~bdd_solver()
    {
        if(lbfgs_) bddb200_lbfgs_destroy(lbfgs_);
        if(solver_) bddb200_destroy(solver_);
    }

    // bdd_solver::read_config, bdd_solver.cpp:468-475: a file name or an inline JSON string
    static json read_config(const std::string& c)
    {
        std::ifstream f(c);
        if(f.good()) { json j; f >> j; return j; }
        return json::parse(c);
    }

    void solve() { solve(config_); }
    void solve(const json& config)
    {
        if(config.contains("export bdd graph")) throw std::runtime_error("'export bdd graph' is not provided by the C++ driver of this build");
        if(solver_ == nullptr)
        {
            read_ILP(config);
            if(config.value("variable order", std::string("input")) != "input") throw std::runtime_error("variable reordering is outside this build's scope");
            if(config.contains("export lp"))
            {   // bdd_solver::export_lp, bdd_solver.cpp:412-430 (the reference also writes .opb and .mps)
                const std::string file = config["export lp"].get<std::string>();
                const size_t dot = file.rfind('.');
                const std::string extension = dot == std::string::npos ? std::string() : file.substr(dot);
                if(extension != ".lp") throw std::runtime_error("Cannot recognize file extension " + extension + " for exporting problem file");
                std::ofstream f(file);
                if(!f.good()) throw std::runtime_error("cannot write " + file);
                f << write_lp(ilp_);
            }
            log("[bdd solver] Compute BDDs");
            bdd_col_ = bdds_from_ilp(ilp_);
            if(config.contains("split bdds"))
            {   // bdd_solver.cpp:105-123 -> bdd_preprocessor.cpp:372-415: "split bdds": {"split length": n}, or a computed length
                const json sb = config["split bdds"].is_object() ? config["split bdds"] : json::object();
                // the reference tests the key "implication bdd" and then reads "implication" (bdd_solver.cpp:119); both spellings are taken here
                const bool implication = sb.value("implication bdd", false) || sb.value("implication", false);
                const size_t length = sb.contains("split length") ? sb["split length"].get<size_t>() : compute_split_length(bdd_col_);
                if(length != std::numeric_limits<size_t>::max())
                {
                    size_t n_split = 0;
                    const size_t before = bdd_col_.nr_bdds();
                    split_long_bdds(bdd_col_, length, ilp_.nr_variables(), implication, &n_split);      // in place; auxiliary variables carry no cost
                    log("[bdd preprocessor] split " + std::to_string(n_split) + " BDDs longer than " + std::to_string(length) + ": " + std::to_string(before) + " -> " + std::to_string(bdd_col_.nr_bdds()) + " BDDs");
                }
            }
            if(config.contains("export bdd lp"))
            {   // bdd_solver::export_bdd_lp, bdd_solver.cpp:400-410
                const std::string file = config["export bdd lp"].get<std::string>();
                std::ofstream f(file);
                if(!f.good()) throw std::runtime_error("cannot write " + file);
                bdd_col_.write_bdd_lp(f, ilp_.objective);
            }
            if(config.contains("print statistics"))
                log("[print_statistics] #variables = " + std::to_string(ilp_.nr_variables()) + ", #constraints = " + std::to_string(ilp_.constraints.size()) + ", #BDDs = " + std::to_string(bdd_col_.nr_bdds()));
            construct_solver(config);
        }
        solve_dual(config);
        perturbation_rounding(config);
    }

    double lower_bound() { require_solver(); double lb = 0; check(bddb200_lower_bound(solver_, &lb)); return lb; }
    // bdd_solver::min_marginals (bdd_solver.cpp:497-514): per variable the (mm_lo, mm_hi) pair of each BDD it occurs in.  The reference
    // throws for its CUDA solvers; the GPU min-marginal sweep of this build answers.
    std::vector<std::vector<std::array<double, 2>>> min_marginals()
    {
        require_solver();
        const size_t nl = bddb200_nr_layers(solver_), nv = bddb200_nr_variables(solver_), nb = bddb200_nr_bdds(solver_);
        std::vector<int32_t> var(nl);
        std::vector<double> lo(nl), hi(nl);
        check(bddb200_min_marginals_host(solver_, 1, var.data(), lo.data(), hi.data()));
        std::vector<std::vector<std::array<double, 2>>> out(nv);
        for(size_t i = 0; i + nb < nl; ++i) out[(size_t)var[i]].push_back({lo[i], hi[i]});       // sorted by variable, the terminal entries last
        return out;
    }
    // export_min_marginals_with_names: (variable name, mm_lo, mm_hi) averaged over the variable's BDDs
    std::tuple<std::vector<std::string>, std::vector<double>, std::vector<double>> min_marginals_with_variable_names()
    {
        const auto mms = min_marginals();
        std::tuple<std::vector<std::string>, std::vector<double>, std::vector<double>> out;
        for(size_t v = 0; v < mms.size() && v < ilp_.nr_variables(); ++v)
        {
            double l = 0, h = 0;
            for(const auto& m : mms[v]) { l += m[0]; h += m[1]; }
            const double n = std::max<size_t>(mms[v].size(), 1);
            std::get<0>(out).push_back(ilp_.var_names[v]); std::get<1>(out).push_back(l / n); std::get<2>(out).push_back(h / n);
        }
        return out;
    }
    double dual_lower_bound() const { return dual_lower_bound_; }     // the bound solve_dual ended with (rounding perturbs the costs afterwards)
    const std::vector<char>& solution() const { return solution_; }
    bool has_solution() const { return solved_; }
    double solution_objective() const
    {
        double obj = ilp_.constant;
        for(size_t i = 0; i < ilp_.nr_variables(); ++i) obj += ilp_.objective[i] * (solution_[i] ? 1.0 : 0.0);
        return obj;
    }
    const ILP& ilp() const { return ilp_; }
    size_t iterations_done() const { return iterations_; }
    bool verbose = true;

private:
    static void check(int code)
    {
        if(code != BDDB200_OK) throw std::runtime_error(std::string("bdd_b200: ") + bddb200_last_error());
    }
    void require_solver() const { if(solver_ == nullptr) throw std::runtime_error("bdd_solver: no solver constructed yet (call solve first)"); }
    void log(const std::string& s) const { if(verbose) std::fprintf(stderr, "%s\n", s.c_str()); }

    // bdd_solver.cpp:44-66
    void read_ILP(const json& config)
    {
        if(!config.contains("input")) throw std::runtime_error("no input specified");
        const std::string inp = config["input"].get<std::string>();
        std::ifstream f(inp);
        if(f.good())
        {
            log("[bdd_solver] Read input file " + inp);
            std::stringstream ss; ss << f.rdbuf();
            ilp_ = parse_lp(ss.str());
        }
        else { log("[bdd_solver] Read input string"); ilp_ = parse_lp(inp); }
    }

    // bdd_solver.cpp:130-267
    void construct_solver(const json& config)
    {
        const std::string precision = config.value("precision", std::string("double"));
        if(precision != "float" && precision != "double") throw std::runtime_error("precision must be float or double");
        const std::string kind = config.value("relaxation solver", std::string("cuda parallel mma"));
        const bool is_mma = kind == "cuda parallel mma";
        const bool is_lbfgs = kind == "lbfgs cuda mma" || kind == "cuda lbfgs parallel mma" || kind == "lbfgs cuda parallel mma";
        if(!is_mma && !is_lbfgs) throw std::runtime_error("solver " + kind + " unknown (this build provides: cuda parallel mma, lbfgs cuda mma, cuda lbfgs parallel mma, lbfgs cuda parallel mma)");
        bddb200_options opt;
        bddb200_default_options(&opt);
        opt.device = config.value("device", 0);
        log("[bdd solver] construct " + std::string(is_lbfgs ? "lbfgs " : "") + "cuda parallel mma solver with " + precision + " precision");
        check(bddb200_create(bdd_col_.instrs.data(), bdd_col_.instrs.size(), bdd_col_.delims.data(), bdd_col_.nr_bdds(), ilp_.objective.data(), ilp_.objective.size(),
                             precision == "double" ? BDDB200_DOUBLE : BDDB200_FLOAT, &opt, &solver_));
        if(is_lbfgs)
        {
            const json l = (kind != "cuda lbfgs parallel mma" && config.contains("lbfgs")) ? config["lbfgs"] : json::object();      // :251-264 takes the defaults only
            check(bddb200_lbfgs_create(solver_, l.value("history size", 5), l.value("initial step size", 1e-6), l.value("required relative lb increase", 1e-6),
                                       l.value("step size decrease factor", 0.8), l.value("step size increase factor", 1.1), &lbfgs_));
        }
    }

    // bdd_solver.cpp:277-309 + include/run_solver_util.h:10-77
    void solve_dual(const json& config)
    {
        const json tc = config.contains("termination criteria") ? config["termination criteria"] : json::object();
        const size_t max_iter = tc.value("maximum iterations", (size_t)1000);
        const double tolerance = tc.value("minimum improvement", 1e-6), improvement_slope = tc.value("improvement slope", 1e-9), time_limit = tc.value("time limit", 3600.0);
        const auto start = std::chrono::steady_clock::now();
        auto spent = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count(); };
        const double lb_initial = lower_bound();
        double lb_first_iter = std::numeric_limits<double>::max(), lb_prev = lb_initial, lb_post = lb_initial;
        log("[bdd solver] initial lower bound = " + num(lb_initial) + ", time = " + num(spent()) + " s");
        for(size_t it = 0; it < max_iter; ++it)
        {
            if(lbfgs_) check(bddb200_lbfgs_iteration(lbfgs_)); else check(bddb200_iteration(solver_, 0.5));
            ++iterations_;
            lb_prev = lb_post;
            lb_post = lower_bound();
            if(it == 0) lb_first_iter = lb_post;
            log("[bdd solver] iteration " + std::to_string(it) + ", lower bound = " + num(lb_post) + ", time = " + num(spent()) + " s");
            if(spent() > time_limit) { log("[bdd solver] Time limit reached."); break; }
            if(std::abs(lb_prev - lb_post) < std::abs(tolerance * lb_prev)) { log("[bdd solver] Relative progress less than tolerance (" + num(tolerance) + ")"); break; }
            if(std::abs(lb_prev - lb_post) < improvement_slope * std::abs(lb_initial - lb_first_iter)) { log("[bdd solver] Improvement slope smaller than " + num(improvement_slope)); break; }
            if(lb_post == std::numeric_limits<double>::infinity()) { log("[bdd solver] problem infeasible"); break; }
        }
        dual_lower_bound_ = lb_post;
        log("[bdd solver] Terminated dual optimization: final lower bound = " + num(lb_post) + ", time = " + num(spent()) + " s");
    }

    // bdd_solver.cpp:318-380
    void perturbation_rounding(const json& config)
    {
        if(!config.contains("perturbation rounding")) return;
        const json pr = config["perturbation rounding"].is_object() ? config["perturbation rounding"] : json::object();
        solution_.assign(bddb200_nr_variables(solver_), 0);
        int solved = 0, rounds = 0;
        check(bddb200_incremental_mm_agreement_rounding(solver_, lbfgs_, pr.value("initial perturbation", 0.1), pr.value("perturbation growth rate", 1.1),
                                                        pr.value("inner iterations", 100), pr.value("outer iterations", 100), solution_.data(), &solved, &rounds));
        solved_ = solved != 0;
        log("[incremental primal rounding] solution objective = " + (solved_ ? num(solution_objective()) : std::string("inf")) + " after " + std::to_string(rounds) + " rounds");
    }

    static std::string num(double x) { char b[64]; std::snprintf(b, sizeof(b), "%.10g", x); return b; }

    json config_;
    ILP ilp_;
    BddCollection bdd_col_;
    bddb200_solver* solver_ = nullptr;
    bddb200_lbfgs* lbfgs_ = nullptr;
    std::vector<char> solution_;
    bool solved_ = false;
    size_t iterations_ = 0;
    double dual_lower_bound_ = -std::numeric_limits<double>::infinity();
};

} // namespace bddb200_host
