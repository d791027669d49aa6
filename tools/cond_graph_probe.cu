// Probe (GPU box): does a CUDA-graph WHILE node accept what one Newton step enqueues -- a cooperative launch, a
// fork / join onto a second stream, device-to-device copies -- and does the device-side cudaGraphSetConditional end
// the loop?   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/cgp tools/cond_graph_probe.cu && /tmp/cgp
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e_ = (x); printf("%-44s %s\n", #x, cudaGetErrorString(e_)); if (e_ != cudaSuccess) ok = 0; } while (0)
__global__ void decide(cudaGraphConditionalHandle h, int* counter, int limit) {
  int c = ++(*counter);
  cudaGraphSetConditional(h, c < limit ? 1u : 0u);
}
__global__ void body(int* x) { atomicAdd(x, 1); }
__global__ void coop(int* x) {
  cg::grid_group g = cg::this_grid();
  if (threadIdx.x == 0) atomicAdd(x, 1);
  g.sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) x[1] = x[0];
}
int main() {
  int ok = 1;
  cudaStream_t s, s2;
  CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
  cudaEvent_t ef, ej;
  cudaEventCreateWithFlags(&ef, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ej, cudaEventDisableTiming);
  int* d;
  cudaMalloc(&d, 64);
  cudaMemset(d, 0, 64);
  cudaGraph_t g;
  CK(cudaGraphCreate(&g, 0));
  cudaGraphConditionalHandle h;
  CK(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams p = {};
  p.type = cudaGraphNodeTypeConditional;
  p.conditional.handle = h;
  p.conditional.type = cudaGraphCondTypeWhile;
  p.conditional.size = 1;
  cudaGraphNode_t node;
  CK(cudaGraphAddNode(&node, g, nullptr, 0, &p));
  cudaGraph_t bodyg = p.conditional.phGraph_out[0];
  CK(cudaStreamBeginCaptureToGraph(s, bodyg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  body<<<4, 32, 0, s>>>(d + 2);
  CK(cudaEventRecord(ef, s));
  CK(cudaStreamWaitEvent(s2, ef, 0));
  body<<<2, 32, 0, s2>>>(d + 3);                        // forked work
  CK(cudaEventRecord(ej, s2));
  int* dp = d + 4;
  void* args[] = {&dp};
  CK(cudaLaunchCooperativeKernel((void*)coop, dim3(64), dim3(128), args, 0, s));
  CK(cudaMemcpyAsync(d + 8, d + 2, 8, cudaMemcpyDeviceToDevice, s));
  CK(cudaStreamWaitEvent(s, ej, 0));                    // join
  decide<<<1, 1, 0, s>>>(h, d, 5);
  cudaGraph_t out;
  CK(cudaStreamEndCapture(s, &out));
  cudaGraphExec_t ex;
  CK(cudaGraphInstantiate(&ex, g, 0));
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(d, 0, 64);
    CK(cudaGraphLaunch(ex, s));
    CK(cudaStreamSynchronize(s));
    int hh[10];
    cudaMemcpy(hh, d, 40, cudaMemcpyDeviceToHost);
    printf("launch %d: iterations %d, body adds %d (expect 5*128=640), forked adds %d (expect 320), coop %d/%d (expect 320)\n", rep, hh[0],
           hh[2], hh[3], hh[4], hh[5]);
    if (hh[0] != 5 || hh[2] != 640 || hh[3] != 320 || hh[4] != 320 || hh[5] != 320) ok = 0;
  }
  printf(ok ? "cond_graph_probe ok\n" : "cond_graph_probe FAILED\n");
  return ok ? 0 : 1;
}
