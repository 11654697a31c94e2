// Standalone hardware check of the patch-extraction kernel (no Python, starts in about a second):
// vg_warp_perspective_u8 on the GPU against the host twin of its per-pixel code (which tests/test_warp_cabi.py holds
// to the cv2-pinned oracle), bit for bit.   build: see tools/build_warp_selftest.sh
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../include/vaegan_b200.h"

static unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

int main() {
  const int sh = 311, sw = 517;
  unsigned seed = 12345u;
  long long total_px = 0, bad = 0;
  int cases = 0;
  for (int ch = 1; ch <= 3; ch += 2) {
    const long long pitch = static_cast<long long>(sw) * ch + 7;
    std::vector<unsigned char> img(static_cast<size_t>(sh) * pitch);
    for (auto& v : img) v = static_cast<unsigned char>(lcg(seed) & 255);
    unsigned char* d_img = nullptr;
    if (cudaMalloc(&d_img, img.size()) != cudaSuccess) { printf("cudaMalloc failed\n"); return 2; }
    cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice);
    const int shapes[4][2] = {{448, 64}, {128, 128}, {256, 256}, {33, 7}};
    for (int s = 0; s < 4; ++s) {
      const int ow = shapes[s][0], oh = shapes[s][1];
      for (int q = 0; q < 6; ++q) {
        float box[8];
        const float cx = 80.f + (lcg(seed) % 350), cy = 60.f + (lcg(seed) % 190), bw = 40.f + (lcg(seed) % 220), bh = 12.f + (lcg(seed) % 90);
        const float base[8] = {cx - bw, cy - bh, cx + bw, cy - bh, cx + bw, cy + bh, cx - bw, cy + bh};
        for (int i = 0; i < 8; ++i) box[i] = base[i] + (static_cast<float>(lcg(seed) % 2000) - 1000.f) * 0.01f * (1 + q);
        double minv[9];
        if (vg_perspective_crop_matrix(box, ow, oh, minv) != 0) { printf("matrix: %s\n", vg_last_error()); continue; }
        const size_t n = static_cast<size_t>(oh) * ow * ch;
        std::vector<unsigned char> h_u8(n), g_u8(n);
        std::vector<float> h_f(n), g_f(n);
        vg_debug_warp_perspective_host(img.data(), sh, sw, ch, pitch, minv, oh, ow, h_u8.data(), h_f.data());
        unsigned char* d_u8 = nullptr; float* d_f = nullptr;
        cudaMalloc(&d_u8, n); cudaMalloc(&d_f, n * sizeof(float));
        cudaMemset(d_u8, 0xAB, n); cudaMemset(d_f, 0xFF, n * sizeof(float));
        const int rc = vg_warp_perspective_u8(d_img, sh, sw, ch, pitch, minv, oh, ow, d_u8, d_f, nullptr);
        if (rc != 0) { printf("launch: %s\n", vg_last_error()); return 3; }
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 4; }
        cudaMemcpy(g_u8.data(), d_u8, n, cudaMemcpyDeviceToHost);
        cudaMemcpy(g_f.data(), d_f, n * sizeof(float), cudaMemcpyDeviceToHost);
        long long b = 0;
        for (size_t i = 0; i < n; ++i) b += (g_u8[i] != h_u8[i]) + (std::memcmp(&g_f[i], &h_f[i], 4) != 0);
        bad += b; total_px += static_cast<long long>(n); ++cases;
        cudaFree(d_u8); cudaFree(d_f);
      }
    }
    cudaFree(d_img);
  }
  int sms = 0, maj = 0, mnr = 0;
  vg_device_info(&sms, &maj, &mnr);
  printf("warp selftest on sm_%d%d (%d SMs): %d cases, %lld values (uint8 + float32 each), mismatches vs host twin: %lld\n", maj, mnr,
         sms, cases, total_px, bad);
  return bad == 0 ? 0 : 1;
}
