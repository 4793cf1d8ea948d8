// tests/cpp/fuzz_layout.cpp -- random BDD collections through the host layout builder (all lane widths, both precisions, shard mode, balanced
// bundles), built with -fsanitize=address,undefined by tests/test_sanitizers.py.  Host code only.
#include <random>
#include <iostream>
#include <numeric>
#include "../../bdd_b200/csrc/host/bdd_solver_native.hpp"
#include "../../bdd_b200/csrc/layout.hpp"
using namespace bddb200_host;
int main()
{
    std::mt19937 rng(11);
    size_t total = 0, errors = 0;
    for(int rep = 0; rep < 80; ++rep)
    {
        ILP ilp;
        const size_t nv = 10 + rng() % 60;
        for(size_t v = 0; v < nv; ++v) { ilp.var_names.push_back("x" + std::to_string(v)); ilp.objective.push_back(1.0); }
        const size_t nc = 1 + rng() % 200;
        for(size_t c = 0; c < nc; ++c)
        {
            Constraint k;
            const size_t len = 1 + rng() % (rep % 4 == 0 ? 30 : 9);
            std::vector<size_t> vars(nv); std::iota(vars.begin(), vars.end(), 0); std::shuffle(vars.begin(), vars.end(), rng); vars.resize(std::min(len, nv)); std::sort(vars.begin(), vars.end());
            long long sum = 0;
            for(size_t v : vars) { k.variables.push_back(v); k.coefficients.push_back(1 + rng() % (rep % 3 == 0 ? 9 : 3)); sum += k.coefficients.back(); }
            k.ineq = rng() % 3; k.rhs = k.ineq == 2 ? (long long)(rng() % 3 + 1) : std::max<long long>(1, sum / 2);
            ilp.constraints.push_back(k);
        }
        BddCollection col;
        try { col = bdds_from_ilp(ilp); } catch(const std::exception&) { continue; }
        if(col.nr_bdds() == 0) continue;
        for(int lanes : {0, 1, 2, 8, 32})
            for(size_t real_bytes : {4, 8})
                for(size_t shared : {(size_t)0, nv / 3})
                {
                    try
                    {
                        const bddb200::HostLayout L = bddb200::build_layout(col.instrs.data(), col.instrs.size(), col.delims.data(), col.nr_bdds(), lanes, nv, real_bytes,
                                                                           bddb200::DEFAULT_STAGE_BUDGET, rep % 2 == 0, rep % 5 == 0 ? 148 : 0, shared);
                        total += L.n_slots + L.n_lay + L.bundles.size();
                    }
                    catch(const bddb200::layout_error&) { ++errors; }
                }
    }
    // malformed collections: children out of range or pointing backwards, missing or doubled sinks, delimiters off by a few, variables
    // shuffled inside a BDD -- the builder has to answer with layout_error (BDDB200_ERR_NOT_QBDD / INVALID_ARGUMENT / TOO_WIDE), not with a fault
    size_t refused = 0, accepted = 0;
    for(int rep = 0; rep < 400; ++rep)
    {
        ILP ilp;
        const size_t nv = 8 + rng() % 12;
        for(size_t v = 0; v < nv; ++v) { ilp.var_names.push_back("x" + std::to_string(v)); ilp.objective.push_back(1.0); }
        for(size_t c = 0; c < 6; ++c)
        {
            Constraint k;
            std::vector<size_t> vars(nv); std::iota(vars.begin(), vars.end(), 0); std::shuffle(vars.begin(), vars.end(), rng); vars.resize(2 + rng() % 5); std::sort(vars.begin(), vars.end());
            for(size_t v : vars) { k.variables.push_back(v); k.coefficients.push_back(1 + rng() % 3); }
            k.ineq = rng() % 3; k.rhs = k.ineq == 2 ? 1 : 2;
            ilp.constraints.push_back(k);
        }
        BddCollection col;
        try { col = bdds_from_ilp(ilp); } catch(const std::exception&) { continue; }
        if(col.nr_bdds() == 0) continue;
        std::vector<bddb200_instruction> ins = col.instrs;
        std::vector<size_t> del = col.delims;
        for(int e = 0, n_edits = 1 + (int)(rng() % 3); e < n_edits; ++e)
        {
            const size_t i = rng() % ins.size();
            switch(rng() % 7)
            {
                case 0: ins[i].lo = rng() % (ins.size() + 4); break;
                case 1: ins[i].hi = rng() % (ins.size() + 4); break;
                case 2: ins[i].index = rng() % 3 == 0 ? (size_t)-1 - rng() % 3 : rng() % (2 * nv); break;
                case 3: ins[i].lo = ins[i].hi = i; break;
                case 4: del[1 + rng() % (del.size() - 1)] += (size_t)(rng() % 5) - 2; break;
                case 5: std::swap(ins[i], ins[rng() % ins.size()]); break;
                default: ins[i].lo = (size_t)-1 - rng() % 4; break;
            }
        }
        try
        {
            const bddb200::HostLayout L = bddb200::build_layout(ins.data(), ins.size(), del.data(), del.size() - 1, 0, 0, 4);
            ++accepted;                                           // an edit can leave a valid collection behind
            total += L.n_slots;
        }
        catch(const bddb200::layout_error&) { ++refused; }
    }
    if(refused < 100) { std::cerr << "only " << refused << " malformed collections refused\n"; return 1; }
    std::cout << "ok " << total << " malformed refused " << refused << " accepted " << accepted << ", layout errors (expected for too-wide BDDs): " << errors << "\n";
    return 0;
}
