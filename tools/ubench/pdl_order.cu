// What does the stream guarantee when a kernel launched with programmatic stream serialization does NOT execute
// griddepcontrol.wait?  A (long) -> B (short, programmatic, no wait) -> {event record | normal kernel C | async copy}.
// Question: can anything ordered after B observe A unfinished?  (The library skips the wait between operations it has
// proven independent; everything that leaves the compute stream must still see all of them complete.)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_order pdl_order.cu && ./pdl_order
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void kA(volatile int* flag, long long spin_clocks) {
    asm volatile("griddepcontrol.launch_dependents;");
    const long long t0 = clock64();
    while (clock64() - t0 < spin_clocks) { }
    if (threadIdx.x == 0 && blockIdx.x == 0) { *flag = 1; __threadfence_system(); }
}
__global__ void kB(volatile int* flagB, volatile const int* flagA, int* sawA) {
    // no griddepcontrol.wait: runs beside A
    if (threadIdx.x == 0) { *sawA = *flagA; *flagB = 1; }
}
__global__ void kC(volatile const int* flagA, int* sawA) {
    if (threadIdx.x == 0) *sawA = *flagA;
}

int main() {
    int *dA, *dB, *dSawB, *dSawC, *hA;
    cudaMalloc(&dA, 4); cudaMalloc(&dB, 4); cudaMalloc(&dSawB, 4); cudaMalloc(&dSawC, 4);
    cudaMallocHost(&hA, 4);
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    int overlapped = 0, ev_early = 0, c_early = 0, copy_early = 0, sync_early = 0;
    const int trials = 200;
    for (int t = 0; t < trials; ++t) {
        cudaMemsetAsync(dA, 0, 4, s); cudaMemsetAsync(dB, 0, 4, s); cudaMemsetAsync(dSawB, 0xFF, 4, s); cudaMemsetAsync(dSawC, 0xFF, 4, s);
        *hA = -1;
        cudaStreamSynchronize(s);
        kA<<<148, 128, 0, s>>>(dA, 400000);   // ~200 us
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(1); cfg.blockDim = dim3(32); cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, kB, (volatile int*)dB, (volatile const int*)dA, dSawB);
        const int mode = t % 4;
        int sawC = 1, hostA = 1;
        if (mode == 0) {            // event record, then look at A's flag from the host through a copy on ANOTHER stream
            cudaEventRecord(e, s);
            cudaEventSynchronize(e);
            cudaMemcpy(&hostA, dA, 4, cudaMemcpyDeviceToHost);   // legacy stream would serialise: use the value as read AFTER the event
            if (hostA != 1) ++ev_early;
        } else if (mode == 1) {     // a normally launched kernel
            kC<<<1, 32, 0, s>>>(dA, dSawC);
            cudaStreamSynchronize(s);
            cudaMemcpy(&sawC, dSawC, 4, cudaMemcpyDeviceToHost);
            if (sawC != 1) ++c_early;
        } else if (mode == 2) {     // an async copy in the same stream
            cudaMemcpyAsync(hA, dA, 4, cudaMemcpyDeviceToHost, s);
            cudaStreamSynchronize(s);
            if (*hA != 1) ++copy_early;
        } else {                    // stream synchronise alone
            cudaStreamSynchronize(s);
            cudaMemcpy(&hostA, dA, 4, cudaMemcpyDeviceToHost);
            if (hostA != 1) ++sync_early;
        }
        cudaDeviceSynchronize();
        int sawB = 1;
        cudaMemcpy(&sawB, dSawB, 4, cudaMemcpyDeviceToHost);
        if (sawB == 0) ++overlapped;   // B really ran while A was unfinished
    }
    printf("trials %d: B overlapped A in %d; A unfinished after [event %d of %d] [normal kernel %d of %d] [async copy %d of %d] [stream sync %d of %d]\n",
           trials, overlapped, ev_early, trials / 4, c_early, trials / 4, copy_early, trials / 4, sync_early, trials / 4);
    printf("%s\n", (ev_early | c_early | copy_early | sync_early) ? "NOT ORDERED: operations after a no-wait kernel can overtake its predecessor"
                                                                 : "ordered: everything after the no-wait kernel saw its predecessor complete");
    return 0;
}
