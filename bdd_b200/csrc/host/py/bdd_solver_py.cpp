// bdd_b200/csrc/host/py/bdd_solver_py.cpp -- pybind11 module `bdd_solver_py`, the Python surface of the JSON-config driver
// (src/bdd_solver/bdd_solver_py.cpp:9-20 of the reference: class bdd_solver with __init__(), __init__(str), __init__(json),
// solve(), solve(json), min_marginals(), min_marginals_with_variable_names(), lower_bound()) over the C++ driver of this build
// (bdd_solver_native.hpp -> C ABI of libbdd_b200.so).  The reference converts Python dicts with pybind11_json.hpp (not in this
// image): here a dict goes through Python's own json.dumps.  Host code only; built by `make -C bdd_b200/csrc py`.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "../bdd_solver_native.hpp"

namespace py = pybind11;
using bddb200_host::bdd_solver;

static nlohmann::json to_json(const py::object& o)
{
    if(py::isinstance<py::str>(o)) return bdd_solver::read_config(o.cast<std::string>());
    return nlohmann::json::parse(py::module_::import("json").attr("dumps")(o).cast<std::string>());
}

PYBIND11_MODULE(bdd_solver_py, m)
{
    m.doc() = "Bindings for the BDD solver driver (cuda parallel mma and its L-BFGS wrapper on B200).";
    py::class_<bdd_solver>(m, "bdd_solver")
        .def(py::init<>())
        .def(py::init([](const py::object& config) { return new bdd_solver(to_json(config)); }), py::arg("config"),
             "config: a JSON file name, an inline JSON string or a dict")
        .def("solve", [](bdd_solver& s) { py::gil_scoped_release nogil; s.solve(); })
        .def("solve", [](bdd_solver& s, const py::object& config) { const nlohmann::json j = to_json(config); py::gil_scoped_release nogil; s.solve(j); })
        .def("min_marginals", &bdd_solver::min_marginals)
        .def("min_marginals_with_variable_names", &bdd_solver::min_marginals_with_variable_names)
        .def("lower_bound", &bdd_solver::lower_bound)
        .def("dual_lower_bound", &bdd_solver::dual_lower_bound)
        .def("has_solution", &bdd_solver::has_solution)
        .def("solution", [](const bdd_solver& s) { std::vector<int> v(s.solution().begin(), s.solution().end()); return v; })
        .def("solution_objective", &bdd_solver::solution_objective)
        .def_readwrite("verbose", &bdd_solver::verbose);
}
