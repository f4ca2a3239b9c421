// sort_yardstick.cu -- YARDSTICK ONLY: CUB DeviceRadixSort::SortPairs and a device-to-device copy beside
// the library's own onesweep sort (pbf_sort_pairs through the C ABI), on the cell keys of the headline scene (8,388,608
// particles, grid 512x256x512, 26 key bits) in the order the simulation sorts them (by particle id).  Not linked into
// libpbf_b200; built by profiles/yardsticks/build.sh into gpurun_out/.
//   usage: sort_yardstick [reps]      prints one JSON line
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <functional>
#include <vector>

#include "pbf_c.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

static float time_ms(cudaStream_t st, int reps, const std::function<void()> &fn) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    fn();
    CK(cudaStreamSynchronize(st));
    cudaEventRecord(a, st);
    for (int i = 0; i < reps; i++) fn();
    cudaEventRecord(b, st);
    CK(cudaStreamSynchronize(st));
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main(int argc, char **argv) {
    const int reps = argc > 1 ? atoi(argv[1]) : 20;
    const int nx = 256, ny = 128, nz = 256, gx = 512, gz = 512;
    const uint32_t n = (uint32_t)nx * ny * nz;
    std::vector<float> pos(4 * (size_t)n), vel(4 * (size_t)n);
    const float origin[3] = {32.5f, 0.5f, 32.5f};
    pbf_scene_dam_break(nx, ny, nz, origin, 0.94f, 0, 12345u, 0u, pos.data(), vel.data());
    std::vector<uint32_t> keys(n), iota(n);
    for (uint32_t i = 0; i < n; i++) {
        const int cx = (int)pos[4 * (size_t)i], cy = (int)pos[4 * (size_t)i + 1], cz = (int)pos[4 * (size_t)i + 2];
        keys[i] = (uint32_t)cx + (uint32_t)cz * gx + (uint32_t)cy * gx * gz;
        iota[i] = i;
    }
    pbf_config cfg = {n, 0, {512, 256, 512}, {16.f, 0.f, 16.f}, 0, -1, 1};
    pbf_handle h;
    if (pbf_create(&cfg, &h)) { fprintf(stderr, "pbf_create: %s\n", pbf_last_error()); return 1; }
    cudaStream_t st = (cudaStream_t)pbf_stream(h);
    uint32_t *kin, *vin, *kout, *vout, *kout2, *vout2;
    CK(cudaMalloc(&kin, 4 * (size_t)n)); CK(cudaMalloc(&vin, 4 * (size_t)n));
    CK(cudaMalloc(&kout, 4 * (size_t)n)); CK(cudaMalloc(&vout, 4 * (size_t)n));
    CK(cudaMalloc(&kout2, 4 * (size_t)n)); CK(cudaMalloc(&vout2, 4 * (size_t)n));
    CK(cudaMemcpy(kin, keys.data(), 4 * (size_t)n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(vin, iota.data(), 4 * (size_t)n, cudaMemcpyHostToDevice));

    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kin, kout, vin, vout, (int)n, 0, 26, st);
    void *tmp;
    CK(cudaMalloc(&tmp, tmp_bytes));
    const float cub_ms = time_ms(st, reps, [&] { cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kin, kout, vin, vout, (int)n, 0, 26, st); });
    const float ours_ms = time_ms(st, reps, [&] {
        if (pbf_sort_pairs(h, kin, vin, kout2, vout2, n, 26)) { fprintf(stderr, "pbf_sort_pairs: %s\n", pbf_last_error()); exit(1); }
    });
    // same result?
    std::vector<uint32_t> a(n), b(n);
    CK(cudaMemcpy(a.data(), vout, 4 * (size_t)n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), vout2, 4 * (size_t)n, cudaMemcpyDeviceToHost));
    size_t diff = 0;
    for (uint32_t i = 0; i < n; i++) diff += a[i] != b[i];
    // streaming yardstick: one pass of a pair sort reads and writes 8 B per element
    const float copy_ms = time_ms(st, reps, [&] {
        cudaMemcpyAsync(kout, kin, 4 * (size_t)n, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(vout, vin, 4 * (size_t)n, cudaMemcpyDeviceToDevice, st);
    });
    // random 26-bit keys: every warp takes the match.any fallback of the run ranking
    srand(1);
    for (uint32_t i = 0; i < n; i++) keys[i] = ((uint32_t)rand() ^ ((uint32_t)rand() << 15)) & ((1u << 26) - 1u);
    CK(cudaMemcpy(kin, keys.data(), 4 * (size_t)n, cudaMemcpyHostToDevice));
    const float cub_rand_ms = time_ms(st, reps, [&] { cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kin, kout, vin, vout, (int)n, 0, 26, st); });
    const float ours_rand_ms = time_ms(st, reps, [&] { pbf_sort_pairs(h, kin, vin, kout2, vout2, n, 26); });
    CK(cudaMemcpy(a.data(), vout, 4 * (size_t)n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), vout2, 4 * (size_t)n, cudaMemcpyDeviceToHost));
    size_t diff_rand = 0;
    for (uint32_t i = 0; i < n; i++) diff_rand += a[i] != b[i];
    printf("{\"n\": %u, \"key_bits\": 26, \"reps\": %d, \"cub_sortpairs_ms\": %.4f, \"pbf_sort_pairs_ms\": %.4f, "
           "\"pbf_passes\": %d, \"d2d_copy_of_one_pass_ms\": %.4f, \"permutation_mismatches_vs_cub\": %zu, "
           "\"random_keys\": {\"cub_sortpairs_ms\": %.4f, \"pbf_sort_pairs_ms\": %.4f, \"permutation_mismatches_vs_cub\": %zu}}\n",
           n, reps, cub_ms, ours_ms, pbf_sort_passes(cfg.grid), copy_ms, diff, cub_rand_ms, ours_rand_ms, diff_rand);
    pbf_destroy(h);
    return 0;
}
