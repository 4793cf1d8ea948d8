// bdd_b200/csrc/host/bdd_solver_cl.cpp -- command-line front end (src/bdd_solver/bdd_solver_cl.cpp:3-10 of the reference):
//     bdd_solver_cl <config.json | inline json>
// reads the JSON configuration, solves the relaxation on the GPU and prints ONE JSON line with the lower bound and, when the
// configuration asks for "perturbation rounding", the primal solution.  With --lp-to-bdds <file.lp> it only converts (no GPU needed)
// and prints the BDD collection's instruction triples (host-side checks of the reader and the BDD builder).
#include <cstdio>
#include <cstring>
#include <iostream>

#include "bdd_solver_native.hpp"

int main(int argc, char** argv)
{
    using namespace bddb200_host;
    try
    {
        if(argc == 3 && std::strcmp(argv[1], "--lp-to-bdds") == 0)
        {
            std::ifstream f(argv[2]);
            if(!f.good()) throw std::runtime_error(std::string("cannot open ") + argv[2]);
            std::stringstream ss; ss << f.rdbuf();
            const ILP ilp = parse_lp(ss.str());
            const BddCollection col = bdds_from_ilp(ilp);
            nlohmann::json out;
            out["nr_variables"] = ilp.nr_variables();
            out["var_names"] = ilp.var_names;
            out["objective"] = ilp.objective;
            out["constant"] = ilp.constant;
            out["delims"] = col.delims;
            std::vector<unsigned long long> flat;
            for(const bddb200_instruction& i : col.instrs) { flat.push_back(i.lo); flat.push_back(i.hi); flat.push_back(i.index); }
            out["instrs"] = flat;
            std::cout << out.dump() << std::endl;
            return 0;
        }
        if((argc == 4 || (argc == 5 && std::strcmp(argv[4], "--implication-bdd") == 0)) && std::strcmp(argv[1], "--split") == 0)
        {   // --split <length | auto> <file.lp> [--implication-bdd]: convert, split the long BDDs, print the collection (host-side check)
            std::ifstream f(argv[3]);
            if(!f.good()) throw std::runtime_error(std::string("cannot open ") + argv[3]);
            std::stringstream ss; ss << f.rdbuf();
            const ILP ilp = parse_lp(ss.str());
            BddCollection col = bdds_from_ilp(ilp);
            const size_t length = std::strcmp(argv[2], "auto") == 0 ? compute_split_length(col) : (size_t)std::stoull(argv[2]);
            size_t n_split = 0, n_vars = ilp.nr_variables();
            if(length != std::numeric_limits<size_t>::max()) n_vars = split_long_bdds(col, length, ilp.nr_variables(), argc == 5, &n_split);
            const std::vector<bddb200_instruction>& ins = col.instrs;
            nlohmann::json out;
            out["split_length"] = length == std::numeric_limits<size_t>::max() ? -1 : (long long)length;
            out["n_split"] = n_split;
            out["nr_variables"] = n_vars;
            out["delims"] = col.delims;
            std::vector<unsigned long long> flat;
            for(const bddb200_instruction& i : ins) { flat.push_back(i.lo); flat.push_back(i.hi); flat.push_back(i.index); }
            out["instrs"] = flat;
            std::cout << out.dump() << std::endl;
            return 0;
        }
        if(argc != 2)
        {
            std::fprintf(stderr, "usage: bdd_solver_cl <config.json | inline json>\n       bdd_solver_cl --lp-to-bdds <file.lp>\n       bdd_solver_cl --split <length | auto> <file.lp> [--implication-bdd]\n");
            return 2;
        }
        bdd_solver s{std::string(argv[1])};
        s.solve();
        nlohmann::json out;
        out["lower_bound"] = s.dual_lower_bound();
        out["iterations"] = s.iterations_done();
        if(s.has_solution())
        {
            nlohmann::json sol = nlohmann::json::object();
            for(size_t i = 0; i < s.ilp().nr_variables(); ++i) sol[s.ilp().var_names[i]] = (int)s.solution()[i];
            out["solution"] = sol;
            out["objective"] = s.solution_objective();
        }
        std::cout << out.dump() << std::endl;
        return 0;
    }
    catch(const std::exception& e)
    {
        std::fprintf(stderr, "bdd_solver_cl: %s\n", e.what());
        return 1;
    }
}
