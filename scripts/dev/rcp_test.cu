// Dev tool: accuracy of the MUFU-seeded reciprocal / rsqrt used by the lane solver, over a wide exponent sweep.
#include <cstdio>
#include <cmath>
#include <cstdint>
#include "../../srb-cbf-nmpc_b200/csrc/lane_prims.h"
__global__ void k(const double* in, double* o1, double* o2, double* s1, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = in[i];
  o1[i] = l_rcp(v); o2[i] = l_rsq(fabs(v));
  double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v)); s1[i] = r;
}
int main() {
  const int n = 1 << 20;
  double *h = new double[n], *a = new double[n], *b = new double[n], *s = new double[n];
  uint64_t st = 88172645463325252ull;
  for (int i = 0; i < n; i++) {
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;
    double m = 1.0 + (double)(st >> 12) / 4503599627370496.0;
    int e = (int)((st >> 3) % 1200) - 600;
    h[i] = ldexp(m, e) * ((st & 1) ? 1 : -1);
  }
  double *d, *d1, *d2, *d3;
  cudaMalloc(&d, n * 8); cudaMalloc(&d1, n * 8); cudaMalloc(&d2, n * 8); cudaMalloc(&d3, n * 8);
  cudaMemcpy(d, h, n * 8, cudaMemcpyHostToDevice);
  k<<<n / 256, 256>>>(d, d1, d2, d3, n);
  cudaMemcpy(a, d1, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(b, d2, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(s, d3, n * 8, cudaMemcpyDeviceToHost);
  double e1 = 0, e2 = 0, e3 = 0; int w1 = 0, w2 = 0;
  for (int i = 0; i < n; i++) {
    double x1 = 1.0 / h[i], x2 = 1.0 / sqrt(fabs(h[i]));
    double r1 = fabs(a[i] - x1) / fabs(x1), r2 = fabs(b[i] - x2) / x2, r3 = fabs(s[i] - x1) / fabs(x1);
    if (!(r1 <= e1)) { e1 = r1; w1 = i; } if (!(r2 <= e2)) { e2 = r2; w2 = i; } if (r3 > e3) e3 = r3;
  }
  printf("max rel err: rcp %.3e (v=%.17g got %.17g) rsq %.3e (v=%.17g got %.17g) seed %.3e ; cuda err %s\n", e1, h[w1], a[w1], e2, h[w2], b[w2], e3,
         cudaGetErrorString(cudaGetLastError()));
  return 0;
}
