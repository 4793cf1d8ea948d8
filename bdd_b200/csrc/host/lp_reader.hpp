// bdd_b200/csrc/host/lp_reader.hpp -- reader for the .lp subset of the reference's PEGTL grammar (src/ILP/ILP_parser.cpp:25-160) and the step
// from an ILP to its BDD collection (bdd_preprocessor::add_ilp, src/bdd_conversion/bdd_preprocessor.cpp:123-228), in plain C++.
// Used by the JSON-config driver (bdd_solver_native.hpp) and exported through the C ABI (include/bdd_b200_collection.h: bddb200_ilp_*).
// Same results as bdd_b200/lp.py + bdd_b200/instances.py (tests/test_cpp_driver.py, tests/test_collection.py), which tests/test_host.py pins
// against the reference's own converter on every fixture.
#pragma once

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "bdd_collection.hpp"

namespace bddb200_host {

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
        {   // digits, optionally a point and more digits (the reference's grammar and bdd_b200/lp.py); anything else is not a right-hand side yet
            size_t q = e;
            while(q < p.size() && std::isdigit((unsigned char)p[q])) ++q;
            const bool has_int = q > e;
            if(q < p.size() && p[q] == '.') { ++q; while(q < p.size() && std::isdigit((unsigned char)p[q])) ++q; }
            if(!has_int || q != p.size()) continue;
        }
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

// ILP_input::feasible / evaluate (include/ILP/ILP_input.h:147-200): every constraint holds; objective value or +infinity
template<typename SOLUTION>
inline bool feasible(const ILP& ilp, const SOLUTION& x)
{
    if(x.size() != ilp.nr_variables()) return false;
    for(const Constraint& c : ilp.constraints)
    {
        long long s = 0;
        for(size_t k = 0; k < c.variables.size(); ++k) s += c.coefficients[k] * (x[c.variables[k]] ? 1 : 0);
        if(c.ineq == LE ? s > c.rhs : (c.ineq == GE ? s < c.rhs : s != c.rhs)) return false;
    }
    return true;
}
template<typename SOLUTION>
inline double evaluate(const ILP& ilp, const SOLUTION& x)
{
    if(!feasible(ilp, x)) return std::numeric_limits<double>::infinity();
    double cost = ilp.constant;
    for(size_t v = 0; v < ilp.objective.size(); ++v) if(x[v]) cost += ilp.objective[v];
    return cost;
}

// one more constraint, terms on the same variable merged, variables in the order they first appear (as the reader does)
inline void add_constraint(ILP& ilp, const std::string& identifier, const std::vector<std::string>& var_names, const std::vector<long long>& coefficients,
                           int ineq, long long rhs)
{
    if(var_names.size() != coefficients.size()) throw std::runtime_error("one coefficient per variable");
    if(ineq != LE && ineq != GE && ineq != EQ) throw std::runtime_error("inequality type not supported");
    Constraint c;
    c.identifier = identifier; c.ineq = ineq; c.rhs = rhs;
    std::map<size_t, long long> merged;
    for(size_t k = 0; k < var_names.size(); ++k)
    {
        const size_t v = ilp.get_or_add_var(var_names[k]);
        if(!merged.count(v)) { merged[v] = 0; c.variables.push_back(v); }
        merged[v] += coefficients[k];
    }
    for(const size_t v : c.variables) c.coefficients.push_back(merged[v]);
    ilp.constraints.push_back(std::move(c));
}

// ILP_input::write_lp (include/ILP/ILP_input.h:229-303): the same sections in the same order; numbers are written with the digits
// they need to read back exactly (the reference uses the stream's default six)
inline std::string write_lp(const ILP& ilp)
{
    std::ostringstream s;
    s.precision(17);
    auto term = [&](double c) { s << (c < 0.0 ? " - " : " + ") << std::abs(c); };
    s << "Minimize\n";
    for(size_t v = 0; v < ilp.objective.size(); ++v) { term(ilp.objective[v]); s << " " << ilp.var_names[v] << "\n"; }
    if(ilp.constant != 0.0) { term(ilp.constant); s << "\n"; }
    s << "Subject To\n";
    for(const Constraint& c : ilp.constraints)
    {
        if(!c.identifier.empty()) s << c.identifier << ":";
        for(size_t k = 0; k < c.variables.size(); ++k)
            if(c.coefficients[k] != 0) { s << (c.coefficients[k] < 0 ? " - " : " + ") << std::llabs(c.coefficients[k]) << " " << ilp.var_names[c.variables[k]]; }
        s << (c.ineq == LE ? " <= " : (c.ineq == GE ? " >= " : " = ")) << c.rhs << "\n";
    }
    s << "Bounds\nBinaries\n";
    for(const std::string& name : ilp.var_names) s << name << "\n";
    s << "End\n";
    return s.str();
}

// bdd_solver::read_ILP (src/bdd_solver/bdd_solver.cpp:44-66): the name of a readable file, else the LP text itself
inline ILP read_ilp(const std::string& file_or_text)
{
    std::ifstream f(file_or_text);
    if(f.good()) { std::stringstream ss; ss << f.rdbuf(); return parse_lp(ss.str()); }
    return parse_lp(file_or_text);
}

} // namespace bddb200_host
