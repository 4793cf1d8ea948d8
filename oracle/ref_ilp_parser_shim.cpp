// oracle/ref_ilp_parser_shim.cpp -- LPMP::ILP_parser / LPMP::OPB_parser for the ref_driver build.  TEST INFRASTRUCTURE ONLY.
//
// The reference reads .lp files with a PEGTL grammar (src/ILP/ILP_parser.cpp) and PEGTL is not in this image, so its own reader cannot
// be compiled.  To LINK the reference's unmodified driver (src/bdd_solver/bdd_solver.cpp, bdd_solver_cl.cpp) against the drop-in GPU
// class, this file supplies the two parser entry points the driver calls (include/ILP/ILP_parser.h, OPB_parser.h): the text goes
// through this repository's reader (bdd_b200/csrc/host/lp_reader.hpp) and is handed to the reference's own ILP_input through its
// public builder interface (include/ILP/ILP_input.h:42-60).  OPB input is refused.
#include <fstream>
#include <sstream>
#include <stdexcept>

#include "ILP/ILP_parser.h"
#include "ILP/OPB_parser.h"
#include "../bdd_b200/csrc/host/lp_reader.hpp"

namespace LPMP {

static ILP_input to_reference(const bddb200_host::ILP& in)
{
    ILP_input out;
    for(size_t v = 0; v < in.nr_variables(); ++v)
    {
        out.add_new_variable(in.var_names[v]);
        out.add_to_objective(in.objective[v], v);
    }
    out.add_to_constant(in.constant);
    for(const bddb200_host::Constraint& c : in.constraints)
    {
        out.begin_new_inequality();
        if(!c.identifier.empty()) out.set_inequality_identifier(c.identifier);
        for(size_t k = 0; k < c.variables.size(); ++k) out.add_to_constraint((int)c.coefficients[k], c.variables[k]);
        out.set_inequality_type(c.ineq == bddb200_host::LE ? ILP_input::inequality_type::smaller_equal
                                : (c.ineq == bddb200_host::GE ? ILP_input::inequality_type::greater_equal : ILP_input::inequality_type::equal));
        out.set_right_hand_side((int)c.rhs);
    }
    return out;
}

namespace ILP_parser {
    ILP_input parse_string(const std::string& input) { return to_reference(bddb200_host::parse_lp(input)); }
    ILP_input parse_file(const std::string& filename)
    {
        std::ifstream f(filename);
        if(!f.good()) throw std::runtime_error("cannot open " + filename);
        std::stringstream ss; ss << f.rdbuf();
        return parse_string(ss.str());
    }
}
namespace OPB_parser {
    ILP_input parse_string(const std::string&) { throw std::runtime_error("OPB input is not available in this build of the reference driver (no PEGTL)"); }
    ILP_input parse_file(const std::string&) { throw std::runtime_error("OPB input is not available in this build of the reference driver (no PEGTL)"); }
}

} // namespace LPMP
