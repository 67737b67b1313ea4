// eikws-b200: lowering of a ModelGraph to the device plan consumed by kernels.cu.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include <cuda_runtime.h>

#include "dev_plan.h"
#include "model_graph.h"

namespace eikws {

// Host-side image of everything that gets uploaded; pointers inside `dev` are offsets into `blob`
// until upload() rebases them onto the device allocation.
struct HostPlan {
    DevPlan dev{};
    std::vector<uint8_t> blob;     // tables, weights, LUTs (16-byte aligned sub-allocations)
    int nn_smem_bytes = 0;         // activation arena + conv row scratch
    std::vector<float> filterbank; // dense [129][32] (debug/parity taps only)
    std::vector<std::pair<size_t, size_t>> fixups;  // (byte offset of a pointer member in DevPlan, offset in blob)
    // debug copies of derived quantities (exposed through eikws_debug_* for parity tests)
};

// Pure host code (no CUDA calls): validates the configuration and computes every derived table.
// Returns 0 or an EIKWS_ERR_* code with a message in err.
int build_host_plan(const ModelGraph &g, HostPlan &hp, std::string &err);

// Host math that mirrors the reference's setup-time arithmetic (exposed for tests)
void quantize_multiplier(double real_multiplier, int32_t *quantized_multiplier, int *shift);

struct DevicePlan {
    DevPlan *d_plan = nullptr;  // device copy of DevPlan
    uint8_t *d_blob = nullptr;
    int nn_smem_bytes = 0;
};
// Uploads to the current CUDA device.
cudaError_t upload_plan(const HostPlan &hp, DevicePlan &dp);
void free_plan(DevicePlan &dp);

}  // namespace eikws
