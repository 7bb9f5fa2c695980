// Developer tool: time ONE instantiation of the fit kernel (model test4, generator path, data
// in shared memory) for compile-time tuning experiments (-DFMC_THREADS=.. -DFMC_MIN_BLOCKS=..).
// Usage: variant_bench <c2.bin> <nres> [reps]     (c2.bin from tools/make_c2_bin.py)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../fitter_mc_b200/csrc/bootstrap_kernel.cuh"
#include "../fitter_mc_b200/csrc/models/models_bench.cuh"

namespace fmc {
#ifndef VARIANT_MODEL
#define VARIANT_MODEL ModelTest4
#endif

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

int run(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 3;
  int n = 0, p = 0;
  if (fread(&n, 4, 1, f) != 1 || fread(&p, 4, 1, f) != 1) return 4;
  std::vector<double> x(n), y(n), err(n), start(p);
  if (fread(x.data(), 8, n, f) != (size_t)n || fread(y.data(), 8, n, f) != (size_t)n ||
      fread(err.data(), 8, n, f) != (size_t)n || fread(start.data(), 8, p, f) != (size_t)p) return 5;
  fclose(f);
  const long long nres = atoll(argv[2]);
  const int reps = argc > 3 ? atoi(argv[3]) : 3;
  const int n_pad = (n + 3) / 4 * 4;
  std::vector<RowData> rows(n_pad);
  for (int i = 0; i < n_pad; ++i) {
    if (i < n) rows[i] = {x[i], y[i], err[i], 1.0 / err[i]};
    else rows[i] = {x[n - 1], 0.0, 0.0, 0.0};
  }
  RowData* drows;
  double *acc, *xbuf, *initb;
  int *istate, *cnt;
  unsigned long long* next;
  constexpr int P = VARIANT_MODEL::NPARAM, Q = VARIANT_MODEL::NPROP, PP = P * (P + 1) / 2;
  CK(cudaMalloc(&drows, sizeof(RowData) * n_pad));
  CK(cudaMalloc(&acc, sizeof(double) * acc_len(P, Q)));
  CK(cudaMalloc(&next, 16));
  CK(cudaMalloc(&xbuf, sizeof(double) * P * nres));
  CK(cudaMalloc(&initb, sizeof(double) * (1 + P + PP) * nres));
  CK(cudaMalloc(&istate, sizeof(int) * 2 * nres));
  CK(cudaMalloc(&cnt, sizeof(int) * 10 * nres));
  CK(cudaMemcpy(drows, rows.data(), sizeof(RowData) * n_pad, cudaMemcpyHostToDevice));
  KernelArgs a{};
  a.rows = drows; a.n = n; a.n_pad = n_pad;
  for (int i = 0; i < P; ++i) a.start[i] = start[i];
  a.seed = 31415; a.nres = nres; a.cap = nres; a.acc = acc;
  a.xbuf = xbuf; a.init = initb; a.istate = istate; a.cnt = cnt;
  size_t shmem = sizeof(RowData) * n_pad;
  auto kinit = init_pass_kernel<VARIANT_MODEL, false, true>;
  auto kiter = iterate_kernel<VARIANT_MODEL, false, true>;
  CK(cudaFuncSetAttribute(kinit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
  CK(cudaFuncSetAttribute(kiter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
  int bps = 0, sms = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kiter, FMC_THREADS, shmem));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  cudaFuncAttributes fa, fb;
  CK(cudaFuncGetAttributes(&fa, kiter));
  CK(cudaFuncGetAttributes(&fb, kinit));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  std::vector<double> hacc(acc_len(P, Q));
  const int ginit = (int)((nres + FMC_THREADS - 1) / FMC_THREADS);
  for (int r = 0; r < reps; ++r) {
    a.first = r * nres;
    CK(cudaMemset(next, 0, 16));
    CK(cudaMemset(acc, 0, sizeof(double) * acc_len(P, Q)));
    cudaEventRecord(e0);
    for (int call = 0; call < 2; ++call) {
      a.call = call;
      a.next_id = next + call;
      kinit<<<ginit, FMC_THREADS, shmem>>>(a);
      kiter<<<sms * bps, FMC_THREADS, shmem>>>(a);
    }
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  CK(cudaMemcpy(hacc.data(), acc, sizeof(double) * hacc.size(), cudaMemcpyDeviceToHost));
  const int NM = 2 * (P + Q);
  printf("threads=%d minblocks=%d regs(iter)=%d local=%zu regs(init)=%d bps=%d grid=%d | nres=%lld best %.3f ms -> %.4e fits/s | nfit %.0f mean0 %.9f passes/fit %.3f\n",
         FMC_THREADS, FMC_MIN_BLOCKS, fa.numRegs, fa.localSizeBytes, fb.numRegs, bps, sms * bps, nres, best, nres / (best * 1e-3),
         hacc[NM + ACC_NFIT], start[0] + hacc[0] / hacc[NM + ACC_NFIT], hacc[NM + ACC_NPASS] / hacc[NM + ACC_NFIT]);
  return 0;
}
}  // namespace fmc

int main(int argc, char** argv) { return fmc::run(argc, argv); }
