// aot_check.cu — ahead-of-time instantiation of the stepper kernels for the built-in
// benchmark right-hand sides.  Build check only (`nvcc -cubin -Xptxas -v`): it shows
// registers/spills of the exact code NVRTC compiles at run time.  Not linked into
// libb200ode.so.
#include "b200_base.cuh"

#if AOT_PROBLEM == 1
// Lorenz, parameters p = (sigma, rho, beta)
// (/root/reference/lib/OrdinaryDiffEqCore/src/precompilation_setup.jl:1-10 form)
__device__ __forceinline__ void aot_rhs(real* du, const real* u, const real* p, const real t) {
    du[0] = p[0] * (u[1] - u[0]);
    du[1] = u[0] * (p[1] - u[2]) - u[1];
    du[2] = u[0] * u[1] - p[2] * u[2];
}
#endif

#define B200_RHS(du, u, p, t) aot_rhs((du), (u), (p), (t))
#include "b200_ensemble.cuh"
