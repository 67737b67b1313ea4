// AddressSanitizer harness for the untrusted-container path (parse_model -> validate_model -> build_host_plan), host code only:
//   g++ -std=c++17 -g -O1 -fsanitize=address,undefined -I include -I ei-keyword-spotting_b200/csrc -I /usr/local/cuda/include \
//       tools/fuzz_container.cpp ei-keyword-spotting_b200/csrc/model_graph.cpp ei-keyword-spotting_b200/csrc/plan.cpp -o fuzz_container
//   ./fuzz_container ei-keyword-spotting_b200/models/*.eikwsmdl          (20,000 random mutations per model)
// plan.cpp's only device-side dependencies are stubbed below.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "model_graph.h"
#include "plan.h"

namespace eikws {
int nn_smem_capacity(bool) { return 8000; }
int nn_smem_capacity_float_graph() { return 19000; }
}  // namespace eikws
extern "C" {
cudaError_t cudaMalloc(void **, size_t) { return cudaErrorUnknown; }
cudaError_t cudaFree(void *) { return cudaSuccess; }
cudaError_t cudaMemcpy(void *, const void *, size_t, cudaMemcpyKind) { return cudaErrorUnknown; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
}

static uint64_t rng_state = 88172645463325252ull;
static uint64_t rnd() {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return rng_state;
}

int main(int argc, char **argv) {
    int refused = 0, accepted = 0;
    for (int a = 1; a < argc; a++) {
        FILE *f = fopen(argv[a], "rb");
        if (!f) return 2;
        std::vector<uint8_t> blob;
        uint8_t buf[4096];
        size_t n;
        while ((n = fread(buf, 1, sizeof(buf), f)) > 0) blob.insert(blob.end(), buf, buf + n);
        fclose(f);
        for (int it = 0; it < 20000; it++) {
            std::vector<uint8_t> b = blob;
            const int edits = 1 + (int)(rnd() % 3);
            for (int e = 0; e < edits; e++) {
                const size_t off = 8 + rnd() % (b.size() - 12);
                const uint32_t choices[] = {0u, 1u, 0xffffffffu, 0x7fffffffu, 0x80000000u, 1u << 20, (uint32_t)rnd(), (uint32_t)(rnd() % 64)};
                const uint32_t v = choices[rnd() % 8];
                memcpy(&b[off & ~size_t(3)], &v, 4);
            }
            if (rnd() % 16 == 0) b.resize(12 + rnd() % (b.size() - 12));
            eikws::ModelGraph g;
            eikws::HostPlan hp;
            std::string err;
            if (!eikws::parse_model(b.data(), b.size(), g, err) || eikws::build_host_plan(g, hp, err) != 0) refused++;
            else accepted++;
        }
    }
    printf("fuzz_container: %d mutated containers refused, %d still valid, no memory error\n", refused, accepted);
    return 0;
}
