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
size_t refcu_nr_bdd_nodes(void* h) { refcu_solver* s = static_cast<refcu_solver*>(h); return s->is_double ? s->d->nr_bdd_nodes() : s->f->nr_bdd_nodes(); }

}
