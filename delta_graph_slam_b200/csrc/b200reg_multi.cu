// b200reg — loop-closure batches over several GPUs of one box, behind the C ABI (include/b200reg.h,
// b200reg_batch_*).
//
// hdl_graph_slam::LoopDetector is ONE C++ object in ONE process [REF include/hdl_graph_slam/loop_detector.hpp:33-187;
// apps/delta_graph_slam_nodelet.cpp:816-824]: a drop-in that spreads its candidate registrations over the GPUs of the
// box cannot ask the caller for one process per GPU.  This translation unit is that single-process form:
//
//   * one b200reg handle per device (own stream, own keyframe cache), driven by one host thread each;
//   * WHOLE TARGETS are dealt to devices (same rule as delta_graph_slam_b200.loop_batch.shard_by_target), so a new
//     keyframe's NDT grid and exact-NN structure is built on exactly one GPU — the multi-GPU form of the
//     setInputTarget hoisted out of the candidate loop [REF loop_detector.hpp:124];
//   * keyframe clouds are registered once (b200reg_batch_cloud_put keeps the caller's pointer: keyframe clouds are
//     immutable for the whole run [REF include/hdl_graph_slam/keyframe.hpp:25-59]) and uploaded lazily to the device
//     that first needs them;
//   * the path shards, so the only exchange is ONE all-gather of the fixed-size result records over NCCL
//     (ncclCommInitAll, one communicator per device, grouped ncclAllGather on the handles' own streams), after
//     which device 0 holds every record and one D2H copy hands them to the host.
//
// NCCL is loaded at run time (dlopen "libnccl.so.2"): the library has no link-time dependency on it, single-GPU
// users never touch it, and inside a process that already carries a NCCL (PyTorch) the same copy is used.
#include <dlfcn.h>
#include <cstdlib>
#include <string.h>

#include <algorithm>
#include <map>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200reg.h"
#include "common.cuh"

// internal entry of b200reg_api.cu: the batch is only enqueued, the records stay on the device
extern "C" int b200reg_internal_align_batch_device(b200reg_handle* h, const b200reg_pair* pairs, size_t n_pairs, int with_fitness, double fitness_max_range, size_t min_slots,
                                                   b200reg_result** d_results);

namespace {

// the few NCCL entry points used, with NCCL 2.x's stable C ABI spelled out (no <nccl.h> needed to build)
typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;
constexpr int kNcclUint8 = 1;  // ncclDataType_t::ncclUint8
struct Nccl {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  bool load(std::string& err) {
    if (lib) return true;
    // B200REG_NCCL_LIB names the library explicitly.  Otherwise the soname: a process that already holds an NCCL
    // (a host that imported torch first gets torch's bundled copy) hands that one back, so there is one NCCL per process.
    // RTLD_LOCAL: a library loaded here must not satisfy somebody else's later symbol look-ups.
    if (const char* path = getenv("B200REG_NCCL_LIB")) lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      if (lib) break;
      lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
    }
    if (!lib) { err = std::string("NCCL is not loadable (dlopen libnccl.so.2): ") + dlerror(); return false; }
#define B200_SYM(field, sym)                                                       \
  field = reinterpret_cast<decltype(field)>(dlsym(lib, sym));                      \
  if (!field) { err = std::string("NCCL lacks ") + sym; lib = nullptr; return false; }
    B200_SYM(CommInitAll, "ncclCommInitAll"); B200_SYM(CommDestroy, "ncclCommDestroy"); B200_SYM(GroupStart, "ncclGroupStart"); B200_SYM(GroupEnd, "ncclGroupEnd");
    B200_SYM(AllGather, "ncclAllGather"); B200_SYM(GetErrorString, "ncclGetErrorString"); B200_SYM(GetVersion, "ncclGetVersion");
#undef B200_SYM
    return true;
  }
};

struct HostCloud {
  const float* xyzw;
  size_t n, stride;
};

}  // namespace

struct b200reg_multi {
  std::vector<b200reg_handle*> h;
  std::vector<int> dev;
  std::vector<cudaStream_t> stream;
  std::vector<ncclComm_t> comm;
  Nccl nccl;
  bool use_nccl = false;
  std::map<long long, HostCloud> clouds;            // registered keyframe clouds (caller's memory)
  std::vector<std::set<long long>> resident;        // ids uploaded to each device's cache
  std::vector<void*> d_recv;                        // per device: n_dev * slot records
  std::vector<size_t> recv_cap;
  unsigned char* pin = nullptr;                     // page-locked landing zone of the gathered records
  size_t pin_cap = 0;
  std::string err;
  // last run (b200reg_batch_get_info)
  std::vector<int> pairs_per_device;
  double gather_ms = 0.0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

namespace {

// Whole targets to devices: targets in order of first appearance, each to the device with the fewest pairs so far
// (ties -> lowest index).  Identical to loop_batch.shard_by_target, so the Python multi-process path and this one
// place every pair on the same rank.
void shard_by_target(const b200reg_pair* pairs, size_t n, int n_dev, std::vector<std::vector<size_t>>& shards) {
  std::map<long long, size_t> first;
  std::vector<std::vector<size_t>> groups;
  for (size_t i = 0; i < n; ++i) {
    auto it = first.find(pairs[i].target_id);
    if (it == first.end()) {
      first[pairs[i].target_id] = groups.size();
      groups.emplace_back();
      groups.back().push_back(i);
    } else {
      groups[it->second].push_back(i);
    }
  }
  shards.assign(n_dev, {});
  std::vector<size_t> load(n_dev, 0);
  for (auto& g : groups) {
    int best = 0;
    for (int d = 1; d < n_dev; ++d)
      if (load[d] < load[best]) best = d;
    shards[best].insert(shards[best].end(), g.begin(), g.end());
    load[best] += g.size();
  }
  for (auto& s : shards) std::sort(s.begin(), s.end());
}

}  // namespace

extern "C" {

int b200reg_batch_create(const b200reg_config* cfg, const int* devices, int n_devices, b200reg_multi** out) {
  if (!cfg || !devices || n_devices < 1 || !out) return B200REG_E_INVALID;
  *out = nullptr;
  for (int a = 0; a < n_devices; ++a)
    for (int b = a + 1; b < n_devices; ++b)
      if (devices[a] == devices[b]) return B200REG_E_INVALID;  // one handle (and one NCCL rank) per GPU
  b200reg_multi* m = new (std::nothrow) b200reg_multi();
  if (!m) return B200REG_E_INVALID;
  auto fail = [&](int rc) {
    for (auto* h : m->h) b200reg_destroy(h);
    delete m;
    return rc;
  };
  for (int d = 0; d < n_devices; ++d) {
    b200reg_config c = *cfg;
    c.device = devices[d];
    b200reg_handle* h = nullptr;
    int rc = b200reg_create(&c, &h);
    if (rc != B200REG_OK) return fail(rc);  // no CPU fallback: a missing device fails the whole object
    m->h.push_back(h);
    m->dev.push_back(devices[d]);
    void* st = nullptr;
    b200reg_get_stream(h, &st);
    m->stream.push_back((cudaStream_t)st);
  }
  m->resident.resize(n_devices);
  m->d_recv.assign(n_devices, nullptr);
  m->recv_cap.assign(n_devices, 0);
  m->pairs_per_device.assign(n_devices, 0);
  // NCCL: required for more than one device (the gather is the one exchange step of the path); with a single device
  // it is used when present, so the same code path runs everywhere
  std::string nerr;
  if (m->nccl.load(nerr)) {
    m->comm.assign(n_devices, nullptr);
    ncclResult_t r = m->nccl.CommInitAll(m->comm.data(), n_devices, m->dev.data());
    if (r != 0) {
      nerr = std::string("ncclCommInitAll: ") + m->nccl.GetErrorString(r);
      m->comm.clear();
    } else {
      m->use_nccl = true;
    }
  }
  if (!m->use_nccl && n_devices > 1) return fail(B200REG_E_CUDA);
  cudaSetDevice(m->dev[0]);
  cudaEventCreate(&m->ev0);
  cudaEventCreate(&m->ev1);
  *out = m;
  return B200REG_OK;
}

int b200reg_batch_destroy(b200reg_multi* m) {
  if (!m) return B200REG_E_INVALID;
  for (size_t d = 0; d < m->h.size(); ++d) {
    cudaSetDevice(m->dev[d]);
    cudaStreamSynchronize(m->stream[d]);
    if (m->d_recv[d]) cudaFree(m->d_recv[d]);
  }
  if (m->use_nccl)
    for (auto c : m->comm)
      if (c) m->nccl.CommDestroy(c);
  cudaSetDevice(m->dev[0]);
  if (m->pin) cudaFreeHost(m->pin);
  if (m->ev0) cudaEventDestroy(m->ev0);
  if (m->ev1) cudaEventDestroy(m->ev1);
  for (auto* h : m->h) b200reg_destroy(h);
  delete m;
  return B200REG_OK;
}

const char* b200reg_batch_last_error(const b200reg_multi* m) { return m ? m->err.c_str() : "null batch object"; }

int b200reg_batch_cloud_put(b200reg_multi* m, int64_t id, const float* xyzw, size_t n, size_t stride_bytes) {
  if (!m || (n && !xyzw)) return B200REG_E_INVALID;
  if (stride_bytes < 12 || (stride_bytes % 4) != 0) { m->err = "stride_bytes must be a multiple of 4 and at least 12"; return B200REG_E_INVALID; }
  auto it = m->clouds.find((long long)id);
  if (it != m->clouds.end()) {
    // a re-registered id replaces the cloud everywhere it is resident
    for (size_t d = 0; d < m->h.size(); ++d)
      if (m->resident[d].erase((long long)id)) b200reg_cloud_drop(m->h[d], id);
  }
  m->clouds[(long long)id] = HostCloud{xyzw, n, stride_bytes};
  return B200REG_OK;
}

int b200reg_batch_cloud_drop(b200reg_multi* m, int64_t id) {
  if (!m) return B200REG_E_INVALID;
  if (!m->clouds.erase((long long)id)) return B200REG_E_INVALID;
  for (size_t d = 0; d < m->h.size(); ++d)
    if (m->resident[d].erase((long long)id)) b200reg_cloud_drop(m->h[d], id);
  return B200REG_OK;
}

int b200reg_batch_run(b200reg_multi* m, const b200reg_pair* pairs, size_t n_pairs, int with_fitness, double fitness_max_range, b200reg_result* results) {
  if (!m || (n_pairs && (!pairs || !results))) return B200REG_E_INVALID;
  if (!n_pairs) return B200REG_OK;
  const int n_dev = (int)m->h.size();
  for (size_t i = 0; i < n_pairs; ++i)
    if (!m->clouds.count(pairs[i].target_id) || !m->clouds.count(pairs[i].source_id)) {
      m->err = "batch_run: pair " + std::to_string(i) + " names a cloud id that was never put";
      return B200REG_E_INVALID;
    }
  std::vector<std::vector<size_t>> shards;
  shard_by_target(pairs, n_pairs, n_dev, shards);
  size_t slot = 0;
  for (auto& s : shards) slot = std::max(slot, s.size());
  const size_t rec = sizeof(b200reg_result);

  // ---- per device, one host thread: upload what the share needs, enqueue the batch, leave the records on the device
  std::vector<int> rc(n_dev, B200REG_OK);
  std::vector<std::string> errs(n_dev);
  std::vector<b200reg_result*> d_res(n_dev, nullptr);
  auto work = [&](int d) {
    b200reg_handle* h = m->h[d];
    const auto& idx = shards[d];
    std::vector<b200reg_pair> local(idx.size());
    for (size_t k = 0; k < idx.size(); ++k) local[k] = pairs[idx[k]];
    for (const auto& p : local)
      for (long long id : {(long long)p.target_id, (long long)p.source_id})
        if (!m->resident[d].count(id)) {
          const HostCloud& c = m->clouds.find(id)->second;
          int r = b200reg_cloud_put(h, id, c.xyzw, c.n, c.stride);
          if (r != B200REG_OK) { rc[d] = r; errs[d] = b200reg_last_error(h); return; }
          m->resident[d].insert(id);
        }
    int r = b200reg_internal_align_batch_device(h, local.data(), local.size(), with_fitness, fitness_max_range, slot, &d_res[d]);
    if (r != B200REG_OK) { rc[d] = r; errs[d] = b200reg_last_error(h); }
  };
  if (n_dev == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int d = 0; d < n_dev; ++d) th.emplace_back(work, d);
    for (auto& t : th) t.join();
  }
  for (int d = 0; d < n_dev; ++d)
    if (rc[d] != B200REG_OK) {
      m->err = "device " + std::to_string(m->dev[d]) + ": " + errs[d];
      for (int e = 0; e < n_dev; ++e) { cudaSetDevice(m->dev[e]); cudaStreamSynchronize(m->stream[e]); }
      return rc[d];
    }
  for (int d = 0; d < n_dev; ++d) m->pairs_per_device[d] = (int)shards[d].size();

  // ---- the one exchange step: all-gather of the result records, then one D2H from device 0
  auto cuda_fail = [&](const char* what, cudaError_t e) {
    m->err = std::string(what) + ": " + cudaGetErrorString(e);
    return B200REG_E_CUDA;
  };
  cudaError_t ce;
  const size_t total = (size_t)n_dev * slot * rec;
  if (m->pin_cap < total) {
    cudaSetDevice(m->dev[0]);
    if (m->pin) cudaFreeHost(m->pin);
    m->pin = nullptr;
    m->pin_cap = 0;
    if ((ce = cudaMallocHost((void**)&m->pin, total + total / 4)) != cudaSuccess) return cuda_fail("cudaMallocHost", ce);
    m->pin_cap = total + total / 4;
  }
  if (m->use_nccl) {
    for (int d = 0; d < n_dev; ++d) {
      if (m->recv_cap[d] >= total) continue;
      cudaSetDevice(m->dev[d]);
      cudaStreamSynchronize(m->stream[d]);
      if (m->d_recv[d]) cudaFree(m->d_recv[d]);
      m->d_recv[d] = nullptr;
      m->recv_cap[d] = 0;
      if ((ce = cudaMalloc(&m->d_recv[d], total + total / 4)) != cudaSuccess) return cuda_fail("cudaMalloc", ce);
      m->recv_cap[d] = total + total / 4;
    }
    cudaSetDevice(m->dev[0]);
    cudaEventRecord(m->ev0, m->stream[0]);
    ncclResult_t nr = m->nccl.GroupStart();
    for (int d = 0; d < n_dev && nr == 0; ++d) nr = m->nccl.AllGather(d_res[d], m->d_recv[d], slot * rec, kNcclUint8, m->comm[d], m->stream[d]);
    const ncclResult_t ne = m->nccl.GroupEnd();
    if (nr == 0) nr = ne;
    if (nr != 0) {
      m->err = std::string("ncclAllGather: ") + m->nccl.GetErrorString(nr);
      return B200REG_E_CUDA;
    }
    cudaSetDevice(m->dev[0]);
    cudaEventRecord(m->ev1, m->stream[0]);
    if ((ce = cudaMemcpyAsync(m->pin, m->d_recv[0], total, cudaMemcpyDeviceToHost, m->stream[0])) != cudaSuccess) return cuda_fail("cudaMemcpyAsync", ce);
  } else {
    // single device without NCCL: the "gather" is the device's own array
    cudaSetDevice(m->dev[0]);
    cudaEventRecord(m->ev0, m->stream[0]);
    cudaEventRecord(m->ev1, m->stream[0]);
    if ((ce = cudaMemcpyAsync(m->pin, d_res[0], slot * rec, cudaMemcpyDeviceToHost, m->stream[0])) != cudaSuccess) return cuda_fail("cudaMemcpyAsync", ce);
  }
  for (int d = 0; d < n_dev; ++d) {
    cudaSetDevice(m->dev[d]);
    if ((ce = cudaStreamSynchronize(m->stream[d])) != cudaSuccess) return cuda_fail("cudaStreamSynchronize", ce);
  }
  float ms = 0.f;
  cudaSetDevice(m->dev[0]);
  if (cudaEventElapsedTime(&ms, m->ev0, m->ev1) == cudaSuccess) m->gather_ms = ms;
  for (int d = 0; d < n_dev; ++d)
    for (size_t k = 0; k < shards[d].size(); ++k) memcpy(&results[shards[d][k]], m->pin + ((size_t)d * slot + k) * rec, rec);
  if (!with_fitness)
    for (size_t i = 0; i < n_pairs; ++i) results[i].fitness = 1.7976931348623157e308;
  return B200REG_OK;
}

int b200reg_batch_get_info(b200reg_multi* m, int* n_devices, int* uses_nccl, int* nccl_version, int* pairs_per_device, double* gather_ms) {
  if (!m) return B200REG_E_INVALID;
  if (n_devices) *n_devices = (int)m->h.size();
  if (uses_nccl) *uses_nccl = m->use_nccl ? 1 : 0;
  if (nccl_version) {
    *nccl_version = 0;
    if (m->use_nccl) m->nccl.GetVersion(nccl_version);
  }
  if (pairs_per_device)
    for (size_t d = 0; d < m->h.size(); ++d) pairs_per_device[d] = m->pairs_per_device[d];
  if (gather_ms) *gather_ms = m->gather_ms;
  return B200REG_OK;
}

}  // extern "C"
