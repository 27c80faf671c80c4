__global__ void k(float* a, float* b, float* c) {
  int i = threadIdx.x;
  c[i] = __fdiv_rn(a[i], __fadd_rn(__fsqrt_rn(b[i]), 1e-8f));
}
