#include <cuda_runtime.h>
#include <cstdio>
__global__ void k_body(int* counter, cudaGraphConditionalHandle h)
{
  if (threadIdx.x == 0) { int v = --(*counter); cudaGraphSetConditional(h, v > 0 ? 1u : 0u); }
}
int main()
{
  cudaStream_t st; cudaStreamCreate(&st);
  int* d; cudaMalloc(&d, 4); int five = 5; cudaMemcpy(d, &five, 4, cudaMemcpyHostToDevice);
  cudaGraph_t g; cudaGraphCreate(&g, 0);
  cudaGraphConditionalHandle h; cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
  cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t node; cudaError_t e = cudaGraphAddNode(&node, g, nullptr, 0, &p);
  printf("add node %d\n", (int)e);
  cudaGraph_t body = p.conditional.phGraph_out[0];
  cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
  k_body<<<1, 32, 0, st>>>(d, h);
  cudaStreamEndCapture(st, nullptr);
  cudaGraphExec_t ex; e = cudaGraphInstantiate(&ex, g, 0); printf("inst %d\n", (int)e);
  cudaGraphLaunch(ex, st); cudaStreamSynchronize(st);
  int out; cudaMemcpy(&out, d, 4, cudaMemcpyDeviceToHost); printf("counter %d\n", out);
  return 0;
}
