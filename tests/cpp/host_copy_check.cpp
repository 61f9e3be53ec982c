// CPU check of the pageable-cloud copy pool (delta_graph_slam_b200/csrc/host_copy.hpp): sizes around the split threshold, two
// concurrent callers (one takes the pool, the other copies alone), and a timing line.  Built and run by tests/test_host_copy.py.
#include "../../delta_graph_slam_b200/csrc/host_copy.hpp"
#include <new>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
int main() {
  const size_t sizes[] = {100, 4096, 262144, 262145, 777777, 2097152, 2097168, 16 << 20};
  for (size_t n : sizes) {
    std::vector<unsigned char> a(n), b(n, 0);
    for (size_t i = 0; i < n; ++i) a[i] = (unsigned char)(i * 2654435761u >> 13);
    b200::host_copy(b.data(), a.data(), n);
    if (memcmp(a.data(), b.data(), n)) { printf("MISMATCH at %zu\n", n); return 1; }
  }
  // concurrent callers: one takes the pool, the other copies alone
  std::vector<unsigned char> a(8 << 20, 7), b1(8 << 20), b2(8 << 20);
  std::thread t1([&] { for (int r = 0; r < 200; ++r) b200::host_copy(b1.data(), a.data(), a.size()); });
  std::thread t2([&] { for (int r = 0; r < 200; ++r) b200::host_copy(b2.data(), a.data(), a.size()); });
  t1.join(); t2.join();
  if (memcmp(a.data(), b1.data(), a.size()) || memcmp(a.data(), b2.data(), a.size())) { printf("MISMATCH concurrent\n"); return 1; }
  for (size_t n : {(size_t)786432, (size_t)2097152}) {
    std::vector<unsigned char> x(n, 1), y(n);
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < 500; ++r) b200::host_copy(y.data(), x.data(), n);
    auto t1 = std::chrono::steady_clock::now();
    for (int r = 0; r < 500; ++r) memcpy(y.data(), x.data(), n);
    auto t2 = std::chrono::steady_clock::now();
    printf("%zu bytes: pool %.1f us, memcpy %.1f us\n", n, std::chrono::duration<double, std::micro>(t1 - t0).count() / 500, std::chrono::duration<double, std::micro>(t2 - t1).count() / 500);
  }
  printf("ok\n");
  return 0;
}
