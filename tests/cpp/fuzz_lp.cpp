// tests/cpp/fuzz_lp.cpp -- mutated .lp texts through the C++ reader (bdd_solver_native.hpp: parse_lp) and the BDD builder: every input either
// parses or raises std::exception; built with -fsanitize=address,undefined by tests/test_sanitizers.py.  Host code only.
#include <fstream>
#include <iostream>
#include <random>
#include <sstream>
#include "../../bdd_b200/csrc/host/bdd_solver_native.hpp"
using namespace bddb200_host;
int main(int argc, char** argv)
{
    std::mt19937 rng(3);
    size_t parsed = 0, rejected = 0;
    const std::string alphabet = " \n\t+-*<=>:0123456789.xyzXe_()[]\\/";
    for(int f = 1; f < argc; ++f)
    {
        std::ifstream in(argv[f]);
        std::stringstream ss; ss << in.rdbuf();
        const std::string base = ss.str();
        for(int rep = 0; rep < 300; ++rep)
        {
            std::string text = base;
            const int n_edits = rep == 0 ? 0 : 1 + (int)(rng() % 6);
            for(int e = 0; e < n_edits && !text.empty(); ++e)
            {
                const size_t pos = rng() % text.size();
                switch(rng() % 4)
                {
                    case 0: text.erase(pos, 1 + rng() % 8); break;
                    case 1: text.insert(pos, 1, alphabet[rng() % alphabet.size()]); break;
                    case 2: text[pos] = alphabet[rng() % alphabet.size()]; break;
                    default: text.insert(pos, text.substr(rng() % text.size(), rng() % 20)); break;
                }
            }
            try
            {
                const ILP ilp = parse_lp(text);
                const BddCollection col = bdds_from_ilp(ilp);
                for(size_t b = 0; b < col.nr_bdds(); ++b) if(!col.is_qbdd(b)) { std::cerr << "builder produced a non-QBDD\n"; return 1; }
                ++parsed;
            }
            catch(const std::exception&) { ++rejected; }
        }
    }
    std::cout << "ok parsed " << parsed << " rejected " << rejected << "\n";
    return parsed > 0 ? 0 : 1;
}
