// Test harness (tests only): exposes the product's fast-math WENO5-Z, compiled as host C++,
// so it can be compared with the oracle without a GPU.
#include "../../parthenon_b200/csrc/weno_fast.cuh"
extern "C" void weno_fast_host(const double *q, long n, double *ql, double *qr) {
  for (long i = 0; i < n; ++i)
    pb2::fastmath::WENO5Z(q[5 * i], q[5 * i + 1], q[5 * i + 2], q[5 * i + 3], q[5 * i + 4], ql[i], qr[i]);
}
extern "C" void linear_fast_host(const double *q, long n, double *ql, double *qr) {
  for (long i = 0; i < n; ++i)
    pb2::fastmath::Linear(q[3 * i], q[3 * i + 1], q[3 * i + 2], ql[i], qr[i]);
}
