// Host build of the device math headers, for CPU-side unit tests only
// (tests/test_detmath.py compiles this with g++ -ffp-contract=off).
#include "b200_detmath.cuh"
extern "C" {
double t_log10_cr(double x) { return b200_log10_cr(x); }
double t_exp10_cr(double x) { return b200_exp10_cr(x); }
double t_fastpower(double x, double y) { return b200_fastpower(x, y); }
float t_fastpower_f(float x, float y) { return b200_fastpower(x, y); }
double t_eps(double x) { return b200_eps(x); }
float t_eps_f(float x) { return b200_eps(x); }
}
extern "C" {
double t_div_const(double a, double b) { return b200_div_const(a, b, 1.0 / b); }
float t_div_const_f(float a, float b) { return b200_div_const(a, b, 1.0f / b); }
}
