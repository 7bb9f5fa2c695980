#include <cuda_runtime.h>
#include <cstdio>
__global__ void body(int* c) { if (threadIdx.x == 0) atomicAdd(c, 1); }
__global__ void ctl(int* c, cudaGraphConditionalHandle h) { cudaGraphSetConditional(h, *c < 10 ? 1u : 0u); }
int main() {
  int* d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
  cudaStream_t st; cudaStreamCreate(&st);
  cudaGraph_t g; cudaGraphCreate(&g, 0);
  cudaGraphConditionalHandle h; 
  printf("handle %d\n", (int)cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams np = {}; np.type = cudaGraphNodeTypeConditional; np.conditional.handle = h; np.conditional.type = cudaGraphCondTypeWhile; np.conditional.size = 1;
  cudaGraphNode_t node; printf("add %d\n", (int)cudaGraphAddNode(&node, g, nullptr, 0, &np));
  cudaGraph_t bg = np.conditional.phGraph_out[0];
  printf("cap %d\n", (int)cudaStreamBeginCaptureToGraph(st, bg, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
  body<<<1, 32, 0, st>>>(d); ctl<<<1, 1, 0, st>>>(d, h);
  printf("end %d\n", (int)cudaStreamEndCapture(st, nullptr));
  cudaGraphExec_t ge; printf("inst %d\n", (int)cudaGraphInstantiate(&ge, g, 0));
  printf("launch %d\n", (int)cudaGraphLaunch(ge, st)); 
  printf("sync %d\n", (int)cudaStreamSynchronize(st));
  int hres; cudaMemcpy(&hres, d, 4, cudaMemcpyDeviceToHost); printf("count %d (expect 10)\n", hres);
  cudaMemset(d, 0, 4); cudaGraphLaunch(ge, st); cudaStreamSynchronize(st); cudaMemcpy(&hres, d, 4, cudaMemcpyDeviceToHost); printf("again %d\n", hres);
}
