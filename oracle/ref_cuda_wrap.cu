// oracle/ref_cuda_wrap.cu -- C wrapper around the UNMODIFIED reference CUDA solver.
//
// TEST / BENCHMARK INFRASTRUCTURE ONLY.  Compiled together with the reference's own bdd_cuda_base.cu and bdd_cuda_parallel_mma.cu
// (where they lie under /root/reference; recipe in oracle/Makefile, cereal replaced by oracle/ref_stubs/cereal) into
// oracle/_ref/libbdd_ref_cuda.so.  It gives bench.py --impl reference_cuda and the tests a same-box GPU baseline: the reference's
// `cuda parallel mma` (bdd_cuda_parallel_mma<REAL>::iteration, src/bdd_solver/bdd_cuda_parallel_mma.cu:142-153: about 6H + 10 thrust /
// kernel launches per iteration on the default stream) next to this repository's kernels.  Nothing under bdd_b200/ links or loads it.
// No reference source text is copied: this file only CALLS the reference's public API.
#include <chrono>
#include <memory>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <thrust/copy.h>
#include <thrust/device_vector.h>
#include <tuple>
#include <type_traits>

#define private public          // fill bdd_instructions / bdd_delimiters verbatim (see oracle/ref_wrap.cpp)
#include "bdd_collection/bdd_collection.h"
#undef private
#include "bdd_solver/bdd_cuda_parallel_mma.h"

using namespace LPMP;

namespace {
struct refcu_solver {
    int is_double;
    std::unique_ptr<bdd_cuda_parallel_mma<double>> d;
    std::unique_ptr<bdd_cuda_parallel_mma<float>> f;
};
template<typename SOLVER>
void solver_costs(SOLVER& x, double* lo_out, double* hi_out, double* mm_out)
{
    const auto costs = x.get_solver_costs();
    const auto& lo = std::get<0>(costs); const auto& hi = std::get<1>(costs); const auto& mm = std::get<2>(costs);
    using REAL = typename std::decay_t<decltype(lo)>::value_type;
    std::vector<REAL> tmp(lo.size());
    thrust::copy(lo.begin(), lo.end(), tmp.begin()); std::copy(tmp.begin(), tmp.end(), lo_out);
    thrust::copy(hi.begin(), hi.end(), tmp.begin()); std::copy(tmp.begin(), tmp.end(), hi_out);
    thrust::copy(mm.begin(), mm.end(), tmp.begin()); std::copy(tmp.begin(), tmp.end(), mm_out);
}
}

extern "C" {

void* refcu_solver_new(const size_t* triples, size_t n_instr, const size_t* delimiters, size_t n_bdds, const double* costs, size_t n_costs, int is_double)
{
    BDD::bdd_collection col;
    col.bdd_instructions.resize(n_instr);
    for(size_t i = 0; i < n_instr; ++i)
    {
        col.bdd_instructions[i].lo = triples[3 * i];
        col.bdd_instructions[i].hi = triples[3 * i + 1];
        col.bdd_instructions[i].index = triples[3 * i + 2];
    }
    col.bdd_delimiters.assign(delimiters, delimiters + n_bdds + 1);
    const std::vector<double> c(costs, costs + n_costs);
    refcu_solver* s = new refcu_solver();
    s->is_double = is_double;
    try
    {
        if(is_double) s->d.reset(new bdd_cuda_parallel_mma<double>(col, c));
        else s->f.reset(new bdd_cuda_parallel_mma<float>(col, c));
    }
    catch(...) { delete s; return nullptr; }
    cudaDeviceSynchronize();
    return s;
}
void refcu_solver_free(void* h) { delete static_cast<refcu_solver*>(h); }

double refcu_lower_bound(void* h)
{
    refcu_solver* s = static_cast<refcu_solver*>(h);
    return s->is_double ? s->d->lower_bound() : s->f->lower_bound();
}
// n iterations back to back; returns the seconds they took (device synchronised on both sides)
double refcu_iterations(void* h, size_t n)
{
    refcu_solver* s = static_cast<refcu_solver*>(h);
    cudaDeviceSynchronize();
    const auto t0 = std::chrono::steady_clock::now();
    for(size_t i = 0; i < n; ++i) { if(s->is_double) s->d->iteration(); else s->f->iteration(); }
    cudaDeviceSynchronize();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
size_t refcu_nr_hops(void* h) { refcu_solver* s = static_cast<refcu_solver*>(h); return s->is_double ? s->d->nr_hops() : s->f->nr_hops(); }
// The reference's per-layer vectors in ITS layer order (hop-sorted, bdd_cuda_base.cu:146-188, :240-285): primal variable and BDD of
// every layer (get_primal_variable_index / get_bdd_index, include/bdd_solver/bdd_cuda_base.h:147-148) and the lo / hi costs and deferred
// min-marginal differences (get_solver_costs, :124-128).  tests/test_layer_order_gpu.py pins reference_layer_order() to them.
size_t refcu_nr_layers(void* h) { refcu_solver* s = static_cast<refcu_solver*>(h); return s->is_double ? s->d->nr_layers() : s->f->nr_layers(); }
void refcu_layer_indices(void* h, int* primal_out, int* bdd_out)
{
    refcu_solver* s = static_cast<refcu_solver*>(h);
    const thrust::device_vector<int> p = s->is_double ? s->d->get_primal_variable_index() : s->f->get_primal_variable_index();
    const thrust::device_vector<int> b = s->is_double ? s->d->get_bdd_index() : s->f->get_bdd_index();
    thrust::copy(p.begin(), p.end(), primal_out);
    thrust::copy(b.begin(), b.end(), bdd_out);
}
void refcu_get_solver_costs(void* h, double* lo_out, double* hi_out, double* mm_out)
{
    refcu_solver* s = static_cast<refcu_solver*>(h);
    if(s->is_double) solver_costs(*s->d, lo_out, hi_out, mm_out); else solver_costs(*s->f, lo_out, hi_out, mm_out);
}
size_t refcu_nr_bdd_nodes(void* h) { refcu_solver* s = static_cast<refcu_solver*>(h); return s->is_double ? s->d->nr_bdd_nodes() : s->f->nr_bdd_nodes(); }

}
