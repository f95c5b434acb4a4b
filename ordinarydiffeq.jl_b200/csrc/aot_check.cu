// aot_check.cu — ahead-of-time instantiation of the stepper kernels for the built-in
// benchmark right-hand sides.  Build check only (`nvcc -cubin -Xptxas -v`): it shows
// registers/spills of the exact code NVRTC compiles at run time.  Not linked into
// libb200ode.so.
#include "b200_base.cuh"

// The right-hand sides are the package's own problem library sources
// (problems_library.py), written to build/aot_problems_gen.inc by build.py.
#define AOT_STR2(x) #x
#define AOT_STR(x) AOT_STR2(x)
#include AOT_STR(AOT_PROBLEM_INC)

#define B200_USER_RHS(du, u, p, t) AOT_RHS_NAME((du), (u), (p), (t))
#ifdef AOT_JAC_NAME
#define B200_JAC(J, u, p, t) AOT_JAC_NAME((J), (u), (p), (t))
#define B200_TGRAD(dT, u, p, t) AOT_TGRAD_NAME((dT), (u), (p), (t))
#endif
#include "b200_ensemble.cuh"
