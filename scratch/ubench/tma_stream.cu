// microbenchmark: how fast can one CTA per SM stream global memory with 1-D bulk TMA copies vs plain LDG.128?
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p){return (uint32_t)__cvta_generic_to_shared(p);}
__device__ __forceinline__ void mbar_init(uint64_t* b,uint32_t c){asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;"::"r"(s32(b)),"r"(c):"memory");}
__device__ __forceinline__ void expect(uint64_t* b,uint32_t n){asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"(s32(b)),"r"(n):"memory");}
__device__ __forceinline__ bool tryw(uint64_t* b,uint32_t ph){uint32_t ok;asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0,1,0,p;}":"=r"(ok):"r"(s32(b)),"r"(ph):"memory");return ok;}
__device__ __forceinline__ void bulk(void* d,const void* s,uint32_t n,uint64_t* b){asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"::"r"(s32(d)),"l"(s),"r"(n),"r"(s32(b)):"memory");}

// each CTA streams `chunks` chunks of `bytes` bytes, chunk c of CTA b at offset (c*gridDim.x + b)*bytes; R slots; `split` copies per chunk
__global__ void tma_stream(const uint8_t* src, int chunks, uint32_t bytes, int R, int split, float* sink){
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) uint64_t full[8];
    if(threadIdx.x==0){for(int i=0;i<8;++i)mbar_init(&full[i],1);asm volatile("fence.mbarrier_init.release.cluster;":::"memory");}
    __syncthreads();
    if(threadIdx.x<32){
        const int lane=threadIdx.x;
        const uint32_t piece=((bytes/split)+15u)&~15u;
        float acc=0.f;
        for(int c=0;c<chunks+R;++c){
            if(c>=R){ // consume chunk c-R
                int cc=c-R; int r=cc%R; while(!tryw(&full[r],(cc/R)&1)){}
                acc+=*(volatile float*)(sm+(size_t)r*bytes+lane*4);
            }
            __syncwarp();
            if(c<chunks){
                int r=c%R;
                if(lane==0)expect(&full[r],bytes);
                __syncwarp();
                if(lane<split){uint32_t off=lane*piece; if(off<bytes){uint32_t len=min(piece,bytes-off);
                    bulk(sm+(size_t)r*bytes+off, src+((size_t)c*gridDim.x+blockIdx.x)*bytes+off, len, &full[r]);}}
            }
            __syncwarp();
        }
        if(acc==123.456f)sink[0]=acc;
    }
}
// LDG.128 streaming: 256 threads, each CTA reads the same chunk pattern, unroll U loads in flight per thread
__global__ void ldg_stream(const float4* src, int chunks, uint32_t bytes, float* sink){
    float4 acc=make_float4(0,0,0,0);
    const int per=bytes/16;
    for(int c=0;c<chunks;++c){
        const float4* p=src+((size_t)c*gridDim.x+blockIdx.x)*per;
        for(int i=threadIdx.x;i<per;i+=blockDim.x){float4 v=__ldg(p+i);acc.x+=v.x;acc.y+=v.y;acc.z+=v.z;acc.w+=v.w;}
    }
    if(acc.x==123.456f)sink[0]=acc.x+acc.y+acc.z+acc.w;
}
int main(){
    size_t total=(size_t)1<<30; uint8_t* d; cudaMalloc(&d,total+ (1<<20)); cudaMemset(d,1,total); float* sink; cudaMalloc(&sink,4);
    cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
    int grid=148;
    cudaFuncSetAttribute(tma_stream,cudaFuncAttributeMaxDynamicSharedMemorySize,200*1024);
    uint32_t sizes[]={2048,4096,15360,16384,32768};
    for(uint32_t bytes: sizes) for(int R: {2,4,6}) for(int split: {1,4}){
        if((size_t)R*bytes>190*1024) continue;
        int chunks=(int)(total/((size_t)grid*bytes));
        tma_stream<<<grid,64,R*bytes>>>(d,chunks,bytes,R,split,sink);
        cudaEventRecord(a); tma_stream<<<grid,64,R*bytes>>>(d,chunks,bytes,R,split,sink); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms,a,b);
        printf("tma bytes=%u R=%d split=%d: %.1f us  %.0f GB/s  err=%s\n",bytes,R,split,ms*1e3,(double)chunks*grid*bytes/ms/1e6,cudaGetErrorString(cudaGetLastError()));
    }
    for(int threads: {256,512,1024}) for(int g: {148,296,592}){
        uint32_t bytes=16384; int chunks=(int)(total/((size_t)g*bytes));
        ldg_stream<<<g,threads>>>((const float4*)d,chunks,bytes,sink);
        cudaEventRecord(a); ldg_stream<<<g,threads>>>((const float4*)d,chunks,bytes,sink); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms,a,b);
        printf("ldg grid=%d threads=%d: %.1f us  %.0f GB/s\n",g,threads,ms*1e3,(double)chunks*g*bytes/ms/1e6);
    }
    return 0;
}
