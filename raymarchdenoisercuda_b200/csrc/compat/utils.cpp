// utils.cpp — printGPUProperties of the compat layer (reference include/utils.h:26, src/utils.cpp:5-16): device
// diagnostics, cold path.
#include "utils.h"

void printGPUProperties() {
    int device = 0;
    cudaDeviceProp prop;
    RMD_CHECK_CUDA(cudaGetDevice(&device));
    RMD_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    std::cout << "Device name: " << prop.name << " (sm_" << prop.major << prop.minor << ", " << prop.multiProcessorCount << " SMs)\n"
              << "Shared memory per block: " << prop.sharedMemPerBlock / 1024.0f << " KB (opt-in " << prop.sharedMemPerBlockOptin / 1024.0f << " KB)\n"
              << "Registers per block: " << prop.regsPerBlock << "\n"
              << "Warp size: " << prop.warpSize << "\n"
              << "Shared memory per multiprocessor: " << prop.sharedMemPerMultiprocessor / 1024.0f << " KB\n"
              << "L2 cache: " << prop.l2CacheSize / (1024.0f * 1024.0f) << " MB\n"
              << std::endl;
}
