// bdd_b200/csrc/host/py/ILP_instance_py.cpp -- pybind11 module `ILP_instance_py`: the Python surface of the ILP input
// (src/ILP/ILP_instance_py.cpp:70-132 of the reference: enum inequality_type, class ILP_instance with nr_constraints, nr_variables,
// evaluate, feasible, objective, get_var_name, get_var_index, add_new_variable_with_obj, add_new_constraint, write_lp; functions read_ILP,
// parse_ILP) over this build's reader (host/lp_reader.hpp).  An ILP_instance is what bdd_cuda_parallel_mma_py.bdd_cuda_parallel_mma is
// constructed from (src/bdd_solver/bdd_cuda_parallel_mma_py.cu:39-44).  Not provided: the Eigen exports (export_constraints,
// node_constraint_incidence_matrix, variable_constraint_bounds), preprocess, read_OPB, read_MRF_UAI.  Host code only.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <fstream>

#include "../lp_reader.hpp"

namespace py = pybind11;
using bddb200_host::ILP;

PYBIND11_MODULE(ILP_instance_py, m)
{
    m.doc() = "Python binding for ILP instance (drop-in surface of the reference's ILP_instance_py)";

    py::enum_<bddb200_host::Ineq>(m, "inequality_type")
        .value("smaller_equal", bddb200_host::LE)
        .value("greater_equal", bddb200_host::GE)
        .value("equal", bddb200_host::EQ)
        .export_values();

    py::class_<ILP>(m, "ILP_instance")
        .def(py::init<>())
        .def("nr_constraints", [](const ILP& ilp) { return ilp.constraints.size(); })
        .def("nr_variables", &ILP::nr_variables)
        .def("evaluate", [](const ILP& ilp, const std::vector<int>& sol) { return bddb200_host::evaluate(ilp, sol); })
        .def("feasible", [](const ILP& ilp, const std::vector<int>& sol) { return bddb200_host::feasible(ilp, sol); })
        .def("objective", [](const ILP& ilp) { return ilp.objective; })
        .def("constant", [](const ILP& ilp) { return ilp.constant; })
        .def("get_var_name", [](const ILP& ilp, size_t index) {
            if(index >= ilp.nr_variables()) throw std::out_of_range("variable index");
            return ilp.var_names[index];
        })
        .def("get_var_index", [](const ILP& ilp, const std::string& name) {
            const auto it = ilp.var_index.find(name);
            if(it == ilp.var_index.end()) throw std::out_of_range("variable " + name + " not present");
            return it->second;
        })
        .def("add_new_variable_with_obj", [](ILP& ilp, const std::string& name, double coefficient) {
            if(ilp.var_index.count(name)) throw std::runtime_error("variable " + name + " exists already");
            const size_t index = ilp.get_or_add_var(name);
            ilp.objective[index] += coefficient;
            return index;
        })
        .def("add_new_constraint", [](ILP& ilp, const std::string& constraint_name, const std::vector<std::string>& var_names, const std::vector<int>& coeffs,
                                      int rhs, bddb200_host::Ineq ineq_type) {
            bddb200_host::add_constraint(ilp, constraint_name, var_names, std::vector<long long>(coeffs.begin(), coeffs.end()), ineq_type, rhs);
        })
        .def("constraint", [](const ILP& ilp, size_t c) {
            if(c >= ilp.constraints.size()) throw std::out_of_range("constraint index");
            const bddb200_host::Constraint& k = ilp.constraints[c];
            return std::make_tuple(k.identifier, k.variables, k.coefficients, (bddb200_host::Ineq)k.ineq, k.rhs);
        }, "(identifier, variables, coefficients, inequality_type, right-hand side)")
        .def("write_lp", [](const ILP& ilp, const std::string& file_path) {
            std::ofstream f(file_path);
            if(!f.good()) throw std::runtime_error("cannot write " + file_path);
            f << bddb200_host::write_lp(ilp);
        })
        .def("__str__", [](const ILP& ilp) { return bddb200_host::write_lp(ilp); });

    m.def("read_ILP", [](const std::string& filename) {
        std::ifstream f(filename);
        if(!f.good()) throw std::runtime_error("cannot open " + filename);
        std::stringstream ss; ss << f.rdbuf();
        return bddb200_host::parse_lp(ss.str());
    });
    m.def("parse_ILP", [](const std::string& instance) { return bddb200_host::parse_lp(instance); });
}
