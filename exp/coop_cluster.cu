#include <cooperative_groups.h>
#include <cstdio>
namespace cg = cooperative_groups;
__global__ void __cluster_dims__(8,1,1) __launch_bounds__(512,1) k(int* out, int* scratch){
  cg::grid_group grid = cg::this_grid();
  cg::cluster_group cl = cg::this_cluster();
  if (threadIdx.x==0) scratch[blockIdx.x] = blockIdx.x;
  grid.sync();
  int s=0; for(int i=0;i<gridDim.x;i++) s+=scratch[i];
  cl.sync();
  if (blockIdx.x==0 && threadIdx.x==0) { out[0]=s; out[1]=cl.num_blocks(); out[2]=gridDim.x; }
  long long t0=clock64();
  for(int i=0;i<100;i++) cl.sync();
  long long t1=clock64();
  for(int i=0;i<100;i++) grid.sync();
  long long t2=clock64();
  if (blockIdx.x==0 && threadIdx.x==0) { out[3]=(int)((t1-t0)/100); out[4]=(int)((t2-t1)/100); }
}
int main(){
  int *out,*scr; cudaMalloc(&out,64); cudaMalloc(&scr,4096); cudaMemset(out,0,64);
  void* args[]={&out,&scr};
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100*1024);
  { cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(144); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = 100*1024; cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1; cfg.attrs = at; cfg.numAttrs = 1; int nc = 0; cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, k, &cfg); printf("max active clusters of 8: %d (%s)\n", nc, cudaGetErrorString(e)); }
  for (int grid : {144, 136, 128, 120, 64, 8}) {
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k, dim3(grid), dim3(512), args, 100*1024, 0);
    cudaError_t e2 = cudaDeviceSynchronize();
    int h[8]; cudaMemcpy(h,out,32,cudaMemcpyDeviceToHost);
    printf("grid %d: launch %s sync %s sum=%d cluster=%d grid=%d clsync=%d cyc gridsync=%d cyc\n", grid, cudaGetErrorString(e), cudaGetErrorString(e2), h[0],h[1],h[2],h[3],h[4]);
    cudaGetLastError();
  }
  return 0;
}
