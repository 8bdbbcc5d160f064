// Host-side microbenchmark of the GP0 command encoder (no GPU): instances per second per thread count.
#include "../../duckstation_b200/csrc/encoder.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
extern "C" uint8_t* psxgen_config(int, uint32_t, uint32_t, uint32_t, size_t*);
int main(int argc, char** argv)
{
  const int n_inst = argc > 1 ? atoi(argv[1]) : 512;
  std::vector<uint8_t*> s(n_inst);
  std::vector<size_t> n(n_inst);
  for (int i = 0; i < n_inst; i++)
    s[i] = psxgen_config(4, i, 2000, 0, &n[i]);
  for (int threads : {1, 2, 4, 8, 16, 32})
  {
    std::vector<psx::Encoder> enc;
    for (int i = 0; i < n_inst; i++)
      enc.emplace_back(256, false);
    std::vector<std::vector<psx::PackedCmd>> outc(n_inst);
    std::vector<std::vector<uint16_t>> outp(n_inst);
    for (int i = 0; i < n_inst; i++)
    {
      outc[i].resize(4096);
      outp[i].resize(80000);
    }
    for (int rep = 0; rep < 2; rep++)
    {
      auto t0 = std::chrono::steady_clock::now();
      std::vector<std::thread> pool;
      for (int t = 0; t < threads; t++)
        pool.emplace_back([&, t]() {
          for (int i = t * n_inst / threads; i < (t + 1) * n_inst / threads; i++)
          {
            enc[i].begin_flush();
            enc[i].set_output(outc[i].data(), 4096, outp[i].data(), 80000);
            size_t used;
            enc[i].add_wire(s[i], n[i], &used);
            enc[i].finish_flush();
          }
        });
      for (auto& th : pool)
        th.join();
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      if (rep == 1)
        printf("threads %2d: %.1f ms for %d instances = %.1f us/instance/thread, %.0f instances/s\n", threads, ms, n_inst,
               ms * 1000 * threads / n_inst, n_inst / ms * 1000);
    }
  }
  return 0;
}
