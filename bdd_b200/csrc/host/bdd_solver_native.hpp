// bdd_b200/csrc/host/bdd_solver_native.hpp -- C++ host side of the JSON-config driver for the GPU solvers.
//
// Mirrors LPMP::bdd_solver (include/bdd_solver/bdd_solver.h:45-103, src/bdd_solver/bdd_solver.cpp) for the relaxation solvers this build
// provides -- "cuda parallel mma" and the three spellings of its L-BFGS wrapper -- over the C ABI of libbdd_b200 (include/bdd_b200.h):
//   read_config (:468-475)  read_ILP (:44-66)  transform_to_BDDs (:112-123)  construct_solver (:130-267)  solve_dual (:277-309, the
//   run_solver loop of include/run_solver_util.h:10-77 with its per-iteration log line)  perturbation_rounding (:318-380)  solve (:477-495).
// The reference reads .lp files with a PEGTL grammar (src/ILP/ILP_parser.cpp:25-160) and turns inequalities into BDDs through its BDD
// manager (src/bdd_conversion, src/bdd_manager); neither library is part of this build, so this header carries a small reader for the
// same LP subset and a direct builder of the quasi-reduced BDD of a linear 0/1 constraint (dynamic programme over partial sums, states
// with equal sub-functions merged bottom-up: the canonical form the reference reaches through lineq_bdd -> bdd_mgr -> reorder -> make_qbdd,
// src/bdd_conversion/bdd_preprocessor.cpp:199-215).  Same code as bdd_b200/lp.py and bdd_b200/instances.py:qbdd_template, in C++.
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

namespace bddb200_host {

// relations LE / GE / EQ: host/bdd_collection.hpp

struct Constraint {
    std::string identifier;
    std::vector<size_t> variables;
    std::vector<long long> coefficients;
    int ineq = LE;
    long long rhs = 0;
};

struct ILP {
    std::vector<double> objective;
    double constant = 0.0;
    std::vector<std::string> var_names;
    std::unordered_map<std::string, size_t> var_index;
    std::vector<Constraint> constraints;

    size_t nr_variables() const { return var_names.size(); }
    size_t get_or_add_var(const std::string& name)
    {
        auto it = var_index.find(name);
        if(it != var_index.end()) return it->second;
        const size_t idx = var_names.size();
        var_index.emplace(name, idx);
        var_names.push_back(name);
        objective.push_back(0.0);
        return idx;
    }
};

// ---------------------------------------------------------------------------------------------------------------- LP reader
namespace detail {

inline bool name_start(char c) { return std::isalpha((unsigned char)c) != 0; }
inline bool name_char(char c)
{
    if(std::isalnum((unsigned char)c)) return true;
    switch(c) { case '_': case '-': case '/': case '(': case ')': case '{': case '}': case ',': case '#': case ';': case '[': case ']': case '.': case '\'': return true; default: return false; }
}
inline void skip_ws(const std::string& s, size_t& p) { while(p < s.size() && std::isspace((unsigned char)s[p])) ++p; }
inline bool read_number(const std::string& s, size_t& p, double& out)
{
    size_t q = p;
    if(q >= s.size() || !std::isdigit((unsigned char)s[q])) return false;
    while(q < s.size() && std::isdigit((unsigned char)s[q])) ++q;
    if(q < s.size() && s[q] == '.') { ++q; while(q < s.size() && std::isdigit((unsigned char)s[q])) ++q; }
    if(q < s.size() && (s[q] == 'e' || s[q] == 'E'))
    {
        size_t r = q + 1;
        if(r < s.size() && (s[r] == '+' || s[r] == '-')) ++r;
        if(r < s.size() && std::isdigit((unsigned char)s[r])) { while(r < s.size() && std::isdigit((unsigned char)s[r])) ++r; q = r; }
    }
    out = std::strtod(s.substr(p, q - p).c_str(), nullptr);
    p = q;
    return true;
}

// "+ 2 x - y + 3" -> [(2, x), (-1, y)] and the trailing constant 3  ([+-] [coef] [*] name, src/ILP/ILP_parser.cpp:52-74)
inline void parse_terms(const std::string& text, std::vector<std::pair<double, std::string>>& terms, double& constant)
{
    size_t p = 0;
    constant = 0.0;
    while(true)
    {
        skip_ws(text, p);
        if(p >= text.size()) break;
        double sign = 1.0;
        bool have_sign = false;
        if(text[p] == '+' || text[p] == '-') { sign = text[p] == '-' ? -1.0 : 1.0; have_sign = true; ++p; skip_ws(text, p); }
        double coef = 1.0;
        const bool have_coef = read_number(text, p, coef);
        skip_ws(text, p);
        if(p < text.size() && text[p] == '*') { ++p; skip_ws(text, p); }
        if(p < text.size() && name_start(text[p]))
        {
            size_t q = p;
            while(q < text.size() && name_char(text[q])) ++q;
            terms.emplace_back(sign * coef, text.substr(p, q - p));
            p = q;
            continue;
        }
        skip_ws(text, p);
        if(have_sign && have_coef && p >= text.size()) { constant = sign * coef; break; }       // trailing constant
        throw std::runtime_error("cannot parse LP expression near: " + text.substr(p, 40));
    }
}

inline std::string lower(std::string s) { for(char& c : s) c = (char)std::tolower((unsigned char)c); return s; }
inline std::string strip(const std::string& s)
{
    size_t a = 0, b = s.size();
    while(a < b && std::isspace((unsigned char)s[a])) ++a;
    while(b > a && std::isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}
inline bool is_section(const std::string& line)
{
    const std::string l = lower(strip(line));
    for(const char* k : {"end", "bounds", "binaries", "generals", "coalesce"})
    {
        const size_t n = std::char_traits<char>::length(k);
        if(l.compare(0, n, k) == 0 && (l.size() == n || !std::isalnum((unsigned char)l[n]))) return true;
    }
    return false;
}

} // namespace detail

// ILP_parser::parse_string (src/ILP/ILP_parser.cpp:25-160): variables are numbered in order of first appearance, the objective first
inline ILP parse_lp(const std::string& text)
{
    std::vector<std::string> lines;
    {
        std::string cur;
        for(char c : text)
        {
            if(c == '\r') continue;
            if(c == '\n') { lines.push_back(cur); cur.clear(); } else cur.push_back(c);
        }
        lines.push_back(cur);
    }
    lines.erase(std::remove_if(lines.begin(), lines.end(), [](const std::string& l) { const std::string s = detail::strip(l); return !s.empty() && s[0] == '\\'; }), lines.end());
    ILP ilp;
    size_t i = 0;
    while(i < lines.size() && detail::strip(lines[i]).empty()) ++i;
    {
        const std::string head = i < lines.size() ? detail::lower(detail::strip(lines[i])) : "";
        if(head != "minimize" && head != "minimise" && head != "min") throw std::runtime_error("LP input must start with 'Minimize'");
        ++i;
    }
    std::string obj;
    auto is_st = [](const std::string& l) { const std::string s = detail::lower(detail::strip(l)); return s == "subject to" || s == "st" || s == "s.t." || s == "such that"; };
    while(i < lines.size() && !is_st(lines[i])) { obj += " " + lines[i]; ++i; }
    if(i >= lines.size()) throw std::runtime_error("missing 'Subject To'");
    ++i;
    {
        std::vector<std::pair<double, std::string>> terms;
        detail::parse_terms(obj, terms, ilp.constant);
        for(const auto& t : terms) ilp.objective[ilp.get_or_add_var(t.second)] += t.first;
    }
    std::string pending;
    for(; i < lines.size(); ++i)
    {
        const std::string& ln = lines[i];
        if(detail::strip(ln).empty()) continue;
        if(pending.empty() && detail::is_section(ln)) break;
        pending += " " + ln;
        // relation + integer right-hand side at the end of the (possibly multi-line) constraint
        std::string p = detail::strip(pending);
        size_t e = p.size();
        while(e > 0 && (std::isdigit((unsigned char)p[e - 1]) || p[e - 1] == '.')) --e;
        if(e == p.size()) continue;                                   // no number at the end yet
        size_t n0 = e;
        while(n0 > 0 && std::isspace((unsigned char)p[n0 - 1])) --n0;
        double sgn = 1.0;
        if(n0 > 0 && (p[n0 - 1] == '+' || p[n0 - 1] == '-')) { sgn = p[n0 - 1] == '-' ? -1.0 : 1.0; --n0; while(n0 > 0 && std::isspace((unsigned char)p[n0 - 1])) --n0; }
        int ineq = -1; size_t rel_begin = n0;
        if(n0 >= 2 && (p.compare(n0 - 2, 2, "<=") == 0 || p.compare(n0 - 2, 2, "=<") == 0)) { ineq = LE; rel_begin = n0 - 2; }
        else if(n0 >= 2 && (p.compare(n0 - 2, 2, ">=") == 0 || p.compare(n0 - 2, 2, "=>") == 0)) { ineq = GE; rel_begin = n0 - 2; }
        else if(n0 >= 1 && p[n0 - 1] == '=') { ineq = EQ; rel_begin = n0 - 1; }
        if(ineq < 0) continue;                                         // the number was a coefficient: the constraint goes on
        const double rhs_val = sgn * std::strtod(p.substr(e).c_str(), nullptr);
        std::string lhs = p.substr(0, rel_begin);
        Constraint c;
        {
            size_t q = 0;
            detail::skip_ws(lhs, q);
            size_t r = q;
            while(r < lhs.size() && !std::isspace((unsigned char)lhs[r]) && lhs[r] != ':') ++r;
            size_t t = r;
            detail::skip_ws(lhs, t);
            if(t < lhs.size() && lhs[t] == ':' && r > q) { c.identifier = lhs.substr(q, r - q); lhs = lhs.substr(t + 1); }
        }
        std::vector<std::pair<double, std::string>> terms;
        double constant = 0.0;
        detail::parse_terms(lhs, terms, constant);
        auto integral = [](double x) { return x == std::floor(x); };
        if(!integral(rhs_val) || !integral(constant)) throw std::runtime_error("constraints must have integer coefficients");
        std::map<size_t, long long> merged;
        for(const auto& t : terms)
        {
            if(!integral(t.first)) throw std::runtime_error("constraints must have integer coefficients");
            const size_t v = ilp.get_or_add_var(t.second);
            if(!merged.count(v)) { merged[v] = 0; c.variables.push_back(v); }
            merged[v] += (long long)t.first;
        }
        for(size_t v : c.variables) c.coefficients.push_back(merged[v]);
        c.ineq = ineq;
        c.rhs = (long long)rhs_val - (long long)constant;
        ilp.constraints.push_back(std::move(c));
        pending.clear();
    }
    return ilp;
}

using BddCollection = bdd_collection;       // host/bdd_collection.hpp: instruction array + delimiters, generators, splitting

// One BDD per constraint, in constraint order (bdd_preprocessor::add_ilp, bdd_preprocessor.cpp:123-228); templates are cached per
// (coefficients, relation, right-hand side)
inline BddCollection bdds_from_ilp(const ILP& ilp)
{
    BddCollection col;
    std::map<std::tuple<std::vector<long long>, int, long long>, QbddTemplate> cache;
    for(const Constraint& c : ilp.constraints)
    {
        const auto key = std::make_tuple(c.coefficients, c.ineq, c.rhs);
        auto it = cache.find(key);
        if(it == cache.end()) it = cache.emplace(key, qbdd_template(c.coefficients, c.ineq, c.rhs)).first;
        col.add_bdd(it->second, c.variables);
    }
    return col;
}

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
        for(const char* key : {"export lp", "export bdd lp", "export bdd graph"})
            if(config.contains(key)) throw std::runtime_error(std::string("'") + key + "' is not provided by the C++ driver of this build");
        if(solver_ == nullptr)
        {
            read_ILP(config);
            if(config.value("variable order", std::string("input")) != "input") throw std::runtime_error("variable reordering is outside this build's scope");
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
