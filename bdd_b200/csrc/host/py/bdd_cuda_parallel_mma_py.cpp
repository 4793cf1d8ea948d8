// bdd_b200/csrc/host/py/bdd_cuda_parallel_mma_py.cpp -- pybind11 module `bdd_cuda_parallel_mma_py`: the Python surface of the solver
// class (src/bdd_solver/bdd_cuda_parallel_mma_py.cu:26-71 of the reference: class bdd_cuda_parallel_mma (double) with pickling,
// __repr__, nr_primal_variables, nr_layers([hop]), nr_hops, nr_bdds, lower_bound, compute_and_set_min_marginal_diff(device pointer))
// over the C ABI of libbdd_b200.so.  As in the reference the constructor takes an ILP_instance (module ILP_instance_py of this
// build, host/py/ILP_instance_py.cpp); the text of an .lp file or its name is accepted as well.  BDDs come from this build's QBDD
// builder (host/bdd_collection.hpp).  iteration / iterations are added so the class is usable on its own.  Host code only.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cuda_runtime_api.h>

#include "../bdd_solver_native.hpp"

namespace py = pybind11;

namespace {

void check(int code)
{
    if(code != BDDB200_OK) throw std::runtime_error(std::string("bdd_b200: ") + bddb200_last_error());
}

struct solver_t {
    bddb200_solver* h = nullptr;
    std::vector<int> cum_layers_per_hop;      // nr_layers(hop), bdd_cuda_base.h:103
    solver_t() {}
    solver_t(const solver_t&) = delete;
    solver_t& operator=(const solver_t&) = delete;
    ~solver_t() { if(h) bddb200_destroy(h); }

    static solver_t* from_lp(const std::string& lp, const std::string& precision, int device)
    {
        return from_ilp(bddb200_host::read_ilp(lp), precision, device);
    }
    static solver_t* from_ilp(const bddb200_host::ILP& ilp, const std::string& precision, int device)
    {
        const bddb200_host::BddCollection col = bddb200_host::bdds_from_ilp(ilp);
        auto s = std::make_unique<solver_t>();
        bddb200_options opt;
        bddb200_default_options(&opt);
        opt.device = device;
        check(bddb200_create(col.instrs.data(), col.instrs.size(), col.delims.data(), col.nr_bdds(), ilp.objective.data(), ilp.objective.size(),
                             precision == "float" ? BDDB200_FLOAT : BDDB200_DOUBLE, &opt, &s->h));
        // layer k of every BDD is hop k; the terminal layer of a BDD of H layers sits at hop H
        std::vector<int> per_hop;
        for(size_t b = 0; b < col.nr_bdds(); ++b)
        {
            size_t hop = 0, prev = col.instrs[col.delims[b]].index;
            for(size_t i = col.delims[b]; i < col.delims[b + 1]; ++i)
            {
                if(col.instrs[i].index >= (size_t)-2) break;
                if(col.instrs[i].index != prev) { ++hop; prev = col.instrs[i].index; }
                if(i == col.delims[b] || col.instrs[i - 1].index != col.instrs[i].index) { if(per_hop.size() <= hop) per_hop.resize(hop + 1, 0); per_hop[hop]++; }
            }
            if(per_hop.size() <= hop + 1) per_hop.resize(hop + 2, 0);
            per_hop[hop + 1]++;
        }
        int c = 0;
        for(int n : per_hop) { c += n; s->cum_layers_per_hop.push_back(c); }
        return s.release();
    }
    py::bytes save() const
    {
        size_t n = 0;
        check(bddb200_save_size(h, &n));
        std::string blob(n + sizeof(uint64_t) * (1 + cum_layers_per_hop.size()), '\0');
        uint64_t k = cum_layers_per_hop.size();
        std::memcpy(&blob[0], &k, 8);
        for(size_t i = 0; i < k; ++i) { uint64_t v = (uint64_t)cum_layers_per_hop[i]; std::memcpy(&blob[8 + 8 * i], &v, 8); }
        check(bddb200_save(h, &blob[8 + 8 * k], n, &n));
        return py::bytes(blob);
    }
    static solver_t* load(const py::bytes& b)
    {
        const std::string blob = b;
        if(blob.size() < 8) throw std::runtime_error("bdd_b200: not a pickled bdd_cuda_parallel_mma");
        uint64_t k = 0;
        std::memcpy(&k, blob.data(), 8);
        if(blob.size() < 8 + 8 * k) throw std::runtime_error("bdd_b200: not a pickled bdd_cuda_parallel_mma");
        auto s = std::make_unique<solver_t>();
        for(size_t i = 0; i < k; ++i) { uint64_t v; std::memcpy(&v, blob.data() + 8 + 8 * i, 8); s->cum_layers_per_hop.push_back((int)v); }
        check(bddb200_load(blob.data() + 8 + 8 * k, blob.size() - 8 - 8 * k, 0, &s->h));
        return s.release();
    }
};

} // namespace

PYBIND11_MODULE(bdd_cuda_parallel_mma_py, m)
{
    m.doc() = "Python binding for the bdd-based solver on B200 (drop-in surface of the reference's bdd_cuda_parallel_mma_py)";
    py::module_::import("ILP_instance_py");          // the ILP_instance type the reference's constructor takes (bdd_cuda_parallel_mma_py.cu:39-44)
    py::class_<solver_t>(m, "bdd_cuda_parallel_mma")
        .def(py::pickle([](const solver_t& s) { return s.save(); }, [](const py::bytes& b) { return solver_t::load(b); }))
        .def(py::init([](const bddb200_host::ILP& ilp, const std::string& precision, int device) { return solver_t::from_ilp(ilp, precision, device); }),
             py::arg("ilp"), py::arg("precision") = "double", py::arg("device") = 0, "ilp: an ILP_instance_py.ILP_instance")
        .def(py::init([](const std::string& lp, const std::string& precision, int device) { return solver_t::from_lp(lp, precision, device); }),
             py::arg("lp"), py::arg("precision") = "double", py::arg("device") = 0, "lp: the text of an .lp file, or its file name")
        .def("__repr__", [](const solver_t& s) {
            return std::string("<bdd_cuda_parallel_mma>: ") + "nr_variables: " + std::to_string(bddb200_nr_variables(s.h)) + ", nr_bdds: " + std::to_string(bddb200_nr_bdds(s.h))
                   + ", nr_layers: " + std::to_string(bddb200_nr_layers(s.h));
        })
        .def("nr_primal_variables", [](const solver_t& s) { return bddb200_nr_variables(s.h); })
        .def("nr_layers", [](const solver_t& s) { return bddb200_nr_layers(s.h); })
        .def("nr_layers", [](const solver_t& s, int hop) {
            if(hop < 0 || (size_t)hop >= s.cum_layers_per_hop.size()) throw std::out_of_range("hop index");
            return hop == 0 ? s.cum_layers_per_hop[0] : s.cum_layers_per_hop[hop] - s.cum_layers_per_hop[hop - 1];
        })
        .def("nr_hops", [](const solver_t& s) { return bddb200_nr_hops(s.h); })
        .def("nr_bdds", [](const solver_t& s) { return bddb200_nr_bdds(s.h); })
        .def("lower_bound", [](solver_t& s) { double lb = 0; check(bddb200_lower_bound(s.h, &lb)); return lb; })
        .def("iteration", [](solver_t& s, double omega) { check(bddb200_iteration(s.h, omega)); }, py::arg("omega") = 0.5)
        .def("iterations", [](solver_t& s, size_t n, double omega) { check(bddb200_iterations(s.h, omega, n)); }, py::arg("n"), py::arg("omega") = 0.5)
        .def("distribute_delta", [](solver_t& s) { check(bddb200_distribute_delta(s.h)); })
        .def("min_marginals", [](solver_t& s) {
            const size_t nl = bddb200_nr_layers(s.h);
            std::vector<int32_t> var(nl); std::vector<double> lo(nl), hi(nl);
            check(bddb200_min_marginals_host(s.h, 0, var.data(), lo.data(), hi.data()));
            return std::make_tuple(var, lo, hi);
        }, "(primal variable index, mm_lo, mm_hi) of every layer in layer order (INT_MAX marks terminal layers)")
        .def("compute_and_set_min_marginal_diff", [](solver_t& s, const long mm_diff_out_ptr) {
            // the reference writes float(mm_hi - mm_lo) of every layer (layer order) to device memory the caller allocated
            // (nr_layers() floats), bdd_cuda_parallel_mma_py.cu:55-69
            const size_t nl = bddb200_nr_layers(s.h);
            std::vector<double> lo(nl), hi(nl);
            check(bddb200_min_marginals_host(s.h, 0, nullptr, lo.data(), hi.data()));
            std::vector<float> diff(nl);
            for(size_t i = 0; i < nl; ++i) diff[i] = (float)(hi[i] - lo[i]);
            if(cudaMemcpy(reinterpret_cast<void*>(mm_diff_out_ptr), diff.data(), nl * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
                throw std::runtime_error("compute_and_set_min_marginal_diff: copy to the given device pointer failed");
        });
}
