// Ceiling of the .dat writer's destination: how fast can NEW file pages be produced on this box's tmpfs (or any dir)?
//   g++ -O2 -pthread -o tools/io_bench tools/io_bench.cpp && tools/io_bench /dev/shm
// For F files x T threads per file, every thread fills its own byte range of a fresh 1 GiB file
//   write : pwrite() of 8 MiB buffers          (page allocation + copy under the inode lock)
//   mmap  : MAP_SHARED + MADV_POPULATE_WRITE + memcpy  (what jne_dat_append_batch_strided_mt does)
//   touch : MAP_SHARED + MADV_POPULATE_WRITE only       (page allocation alone: the floor of any writer)
// and, for reference, memcpy into already-resident pages (rewrite).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <sys/mman.h>
#include <thread>
#include <unistd.h>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
  const std::string dir = argc > 1 ? argv[1] : "/dev/shm";
  const size_t fsz = (argc > 2 ? atoll(argv[2]) : 1024) << 20;
  std::vector<unsigned char> src(8 << 20);
  for (size_t i = 0; i < src.size(); ++i) src[i] = (unsigned char)(i * 131);
  printf("# dir %s, %zu MiB per file, hardware threads %u\n", dir.c_str(), fsz >> 20, std::thread::hardware_concurrency());
  printf("%-8s %5s %7s %10s\n", "mode", "files", "threads", "GB/s");
  for (const char* mode : {"write", "mmap", "touch", "rewrite"})
    for (int F : {1, 5})
      for (int T : {1, 2, 4, 8}) {
        if (F * T > 40) continue;
        std::vector<int> fds(F);
        std::vector<std::string> names(F);
        for (int f = 0; f < F; ++f) {
          names[f] = dir + "/jne_io_bench_" + std::to_string(getpid()) + "_" + std::to_string(f);
          fds[f] = open(names[f].c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
          if (fds[f] < 0) { perror("open"); return 1; }
          if (!strcmp(mode, "rewrite")) {   // make the pages resident first
            if (ftruncate(fds[f], fsz)) return 1;
            void* m = mmap(nullptr, fsz, PROT_READ | PROT_WRITE, MAP_SHARED, fds[f], 0);
            memset(m, 1, fsz); munmap(m, fsz);
          }
        }
        const double t0 = now();
        std::vector<std::thread> th;
        for (int f = 0; f < F; ++f) {
          if (strcmp(mode, "write") && ftruncate(fds[f], fsz)) return 1;
          for (int t = 0; t < T; ++t)
            th.emplace_back([&, f, t]() {
              const size_t a = fsz / T * t, b = t == T - 1 ? fsz : fsz / T * (t + 1);
              if (!strcmp(mode, "write")) {
                for (size_t o = a; o < b; o += src.size()) {
                  const size_t k = std::min(src.size(), b - o);
                  if (pwrite(fds[f], src.data(), k, (off_t)o) != (ssize_t)k) { perror("pwrite"); exit(1); }
                }
              } else {
                unsigned char* m = (unsigned char*)mmap(nullptr, b - a, PROT_READ | PROT_WRITE, MAP_SHARED, fds[f], (off_t)a);
                if (m == MAP_FAILED) { perror("mmap"); exit(1); }
#ifdef MADV_POPULATE_WRITE
                if (strcmp(mode, "rewrite")) madvise(m, b - a, MADV_POPULATE_WRITE);
#endif
                if (strcmp(mode, "touch"))
                  for (size_t o = 0; o < b - a; o += src.size()) memcpy(m + o, src.data(), std::min(src.size(), b - a - o));
                munmap(m, b - a);
              }
            });
        }
        for (auto& x : th) x.join();
        const double dt = now() - t0;
        printf("%-8s %5d %7d %10.2f\n", mode, F, T, (double)F * fsz / dt / 1e9);
        fflush(stdout);
        for (int f = 0; f < F; ++f) { close(fds[f]); unlink(names[f].c_str()); }
      }
  return 0;
}
