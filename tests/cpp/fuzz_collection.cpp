// tests/cpp/fuzz_collection.cpp -- random constraints through the host-side BDD collection (split with implication BDDs, bdd_and, reorder,
// make_qbdd, remove, the generators), built with -fsanitize=address,undefined by tests/test_sanitizers.py.  Host code only.
#include <random>
#include <iostream>
#include "../../bdd_b200/csrc/host/bdd_solver_native.hpp"
using namespace bddb200_host;
int main()
{
    std::mt19937 rng(7);
    size_t total = 0;
    for(int rep = 0; rep < 60; ++rep)
    {
        ILP ilp;
        const size_t nv = 12 + rng() % 20;
        for(size_t v = 0; v < nv; ++v) { ilp.var_names.push_back("x" + std::to_string(v)); ilp.objective.push_back(double(rng() % 9) - 4); }
        const size_t nc = 3 + rng() % 8;
        for(size_t c = 0; c < nc; ++c)
        {
            Constraint k;
            const size_t len = 2 + rng() % 11;
            std::vector<size_t> vars(nv); std::iota(vars.begin(), vars.end(), 0); std::shuffle(vars.begin(), vars.end(), rng); vars.resize(std::min(len, nv)); std::sort(vars.begin(), vars.end());
            long long sum = 0;
            for(size_t v : vars) { k.variables.push_back(v); k.coefficients.push_back(1 + rng() % 4); sum += k.coefficients.back(); }
            k.ineq = rng() % 3; k.rhs = k.ineq == 2 ? (long long)(rng() % 3 + 1) : std::max<long long>(1, sum / 2);
            ilp.constraints.push_back(k);
        }
        BddCollection col;
        try { col = bdds_from_ilp(ilp); } catch(const std::exception& e) { continue; }
        if(col.nr_bdds() == 0) continue;
        for(size_t length : {2, 3, 5})
        {
            BddCollection c = col;
            size_t n_split = 0;
            const size_t nvars = split_long_bdds(c, length, nv, true, &n_split);
            for(size_t b = 0; b < c.nr_bdds(); ++b) if(!c.is_qbdd(b) || !c.is_reordered(b)) { std::cerr << "not a layered qbdd\n"; return 1; }
            total += c.nr_bdds() + nvars;
            // and / make_qbdd / reorder / remove on what is there
            if(c.nr_bdds() >= 3)
            {
                std::vector<size_t> some{0, c.nr_bdds() / 2, c.nr_bdds() - 1};
                try { const size_t a = c.bdd_and(some.begin(), some.end()); c.reorder(a); const size_t q = c.make_qbdd(a); if(!c.is_qbdd(q)) return 2; c.remove(a); } catch(const std::invalid_argument&) {}
                c.remove(some.begin(), some.begin() + 2);
            }
        }
    }
    // malformed arrays: the constructor refuses them (children out of range or pointing backwards, sinks missing, delimiters off) or what
    // it lets through is safe to work on
    size_t refused = 0;
    for(int rep = 0; rep < 2000; ++rep)
    {
        bdd_collection good;
        const size_t a = good.cardinality_constraint(3 + rng() % 6, 1 + rng() % 3), b = good.simplex_constraint(2 + rng() % 7);
        (void)a; (void)b;
        std::vector<bddb200_instruction> ins = good.instrs;
        std::vector<size_t> del = good.delims;
        for(int e = 0, n_edits = 1 + (int)(rng() % 3); e < n_edits; ++e)
        {
            const size_t i = rng() % ins.size();
            switch(rng() % 6)
            {
                case 0: ins[i].lo = rng() % (ins.size() + 4); break;
                case 1: ins[i].hi = rng() % (ins.size() + 4); break;
                case 2: ins[i].index = (size_t)-1 - rng() % 3; break;
                case 3: ins[i].lo = ins[i].hi = i; break;
                case 4: del[1 + rng() % (del.size() - 1)] += (size_t)(rng() % 5) - 2; break;
                default: ins[i].hi = (size_t)-1 - rng() % 4; break;
            }
        }
        try
        {
            bdd_collection c(ins.data(), ins.size(), del.data(), del.size() - 1);
            for(size_t k = 0, n_given = c.nr_bdds(); k < n_given; ++k)
            {
                total += c.variables(k).size() + c.is_qbdd(k) + c.is_reordered(k);
                try { c.evaluate(k, std::vector<char>(64, (char)(rep & 1))); } catch(const std::out_of_range&) {}
                try { c.reorder(k); const size_t q = c.make_qbdd(k); total += c.nr_bdd_nodes(q); } catch(const std::exception&) {}
            }
            try { total += c.bdd_and(0, 1); } catch(const std::exception&) {}
        }
        catch(const std::invalid_argument&) { ++refused; }
    }
    if(refused < 500) { std::cerr << "only " << refused << " malformed collections refused\n"; return 1; }
    bdd_collection g;
    for(size_t n = 1; n < 30; ++n) { g.simplex_constraint(n); g.not_all_false_constraint(n); if(n > 1) { g.all_equal_constraint(n); for(size_t k = 0; k <= n; ++k) g.cardinality_constraint(n, k); } }
    std::cout << "ok " << total << " " << g.nr_bdds() << "\n";
    return 0;
}
