// b200reg — loop-closure batches behind the C ABI: the keyframe cloud cache, b200reg_align_batch / calc_fitness_batch
// [REF include/hdl_graph_slam/loop_detector.hpp:119-173; src/hdl_graph_slam/information_matrix_calculator.cpp:53-108].
// Included by b200reg_api.cu (one translation unit: the entry points share the handle definition and the helpers above).
// ---- loop-closure batches --------------------------------------------------------------------
static CachedCloud* cache_find(b200reg_handle* h, long long id) {
  auto it = h->cache.find(id);
  return it == h->cache.end() ? nullptr : &it->second;
}

int b200reg_cloud_put(b200reg_handle* h, int64_t id, const float* xyzw, size_t n, size_t stride) {
  if (!h || (n && !xyzw)) return B200REG_E_INVALID;
  int rc = set_device(h);
  if (rc) return rc;
  CachedCloud& c = h->cache[(long long)id];
  if ((rc = upload_cloud(h, xyzw, n, stride, c.pts, /*defer_pinned=*/true))) return rc;
  c.n = (int)n;
  c.has_ndt = c.has_nn = false;
  return B200REG_OK;
}

int b200reg_cloud_put_device(b200reg_handle* h, int64_t id, const float* d_xyzw, size_t n) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h || (n && !d_xyzw)) return B200REG_E_INVALID;
  int rc = set_device(h);
  if (rc) return rc;
  CachedCloud& c = h->cache[(long long)id];
  B200_CUDA_TRY(c.pts.reserve(n ? n : 1));
  if (n) B200_CUDA_TRY(cudaMemcpyAsync(c.pts.p, d_xyzw, n * 16, cudaMemcpyDeviceToDevice, h->stream));
  c.n = (int)n;
  c.has_ndt = c.has_nn = false;
  return B200REG_OK;
}

int b200reg_cloud_drop(b200reg_handle* h, int64_t id) {
  if (!h) return B200REG_E_INVALID;
  auto it = h->cache.find((long long)id);
  if (it == h->cache.end()) return B200REG_E_INVALID;
  cudaSetDevice(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  it->second.release();
  h->cache.erase(it);
  return B200REG_OK;
}

int b200reg_cloud_clear(b200reg_handle* h) {
  if (!h) return B200REG_E_INVALID;
  cudaSetDevice(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  for (auto& kv : h->cache) kv.second.release();
  h->cache.clear();
  return B200REG_OK;
}

int b200reg_cloud_sync(b200reg_handle* h) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h) return B200REG_E_INVALID;
  int rc = set_device(h);
  if (rc) return rc;
  B200_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return B200REG_OK;
}

int b200reg_cloud_count(b200reg_handle* h, size_t* out) {
  if (!h || !out) return B200REG_E_INVALID;
  *out = h->cache.size();
  return B200REG_OK;
}

// do_align = 0: no registration, the pair's `guess` IS the transform the fitness is evaluated at
// (b200reg_calc_fitness_batch); only the exact-NN product of the targets is needed then
// results == nullptr with d_results_out != nullptr: the records stay on the device (slot i of *d_results_out belongs to
// pairs[i]); everything is only ENQUEUED on the handle's stream — the multi-GPU entry (b200reg_multi.cu) gathers them
// with one NCCL all-gather and synchronises once.  `min_slots`: capacity the record array must have (the gather's slot).
static int batch_run(b200reg_handle* h, const b200reg_pair* pairs, size_t n_pairs, int do_align, int with_fitness, double fitness_max_range, b200reg_result* results,
                     b200reg_result** d_results_out = nullptr, size_t min_slots = 0) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  const bool keep_on_device = results == nullptr && d_results_out != nullptr;
  if (!h || (n_pairs && (!pairs || (!results && !keep_on_device)))) return B200REG_E_INVALID;
  if (keep_on_device) {
    if (set_device(h)) return B200REG_E_CUDA;
    B200_CUDA_TRY(h->batch_results.reserve(std::max(n_pairs, min_slots) + 1));
    *d_results_out = h->batch_results.p;
  }
  if (!n_pairs) return B200REG_OK;
  if (do_align && h->cfg.method != B200REG_METHOD_NDT) { h->err = "align_batch: this handle's registration method has no batch path"; return B200REG_E_STATE; }
  int rc = set_device(h);
  if (rc) return rc;
  const float res = (float)h->cfg.resolution;
  // ---- look up the clouds; build the target products once per distinct target
  std::vector<CachedCloud*> tgt(n_pairs), src(n_pairs);
  for (size_t i = 0; i < n_pairs; ++i) {
    tgt[i] = cache_find(h, pairs[i].target_id);
    src[i] = cache_find(h, pairs[i].source_id);
    if (!tgt[i] || !src[i]) { h->err = "align_batch: pair " + std::to_string(i) + " names a cloud id that was never put"; return B200REG_E_INVALID; }
    if (tgt[i]->n == 0) { h->err = "align_batch: pair " + std::to_string(i) + ": Invalid or empty point cloud dataset given!"; return B200REG_E_INVALID; }
  }
  {
    // fork: the lanes start after everything already queued on the handle's stream (cloud uploads)
    std::vector<CachedCloud*> todo;
    for (size_t i = 0; i < n_pairs; ++i) {
      CachedCloud& c = *tgt[i];
      const bool need = (do_align && (!c.has_ndt || c.ndt_res != res)) || (with_fitness && !c.has_nn);
      if (need && std::find(todo.begin(), todo.end(), &c) == todo.end()) todo.push_back(&c);
    }
    if (!todo.empty()) {
      if (!h->ev_fork) B200_CUDA_TRY(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
      B200_CUDA_TRY(cudaEventRecord(h->ev_fork, h->stream));
      const int n_lanes = (int)std::min<size_t>(todo.size(), (size_t)b200reg_handle::kBuildLanes);
      for (int l = 0; l < n_lanes; ++l) {
        auto& ln = h->lanes[l];
        if (!ln.st) {
          B200_CUDA_TRY(cudaStreamCreateWithFlags(&ln.st, cudaStreamNonBlocking));
          B200_CUDA_TRY(cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming));
        }
        B200_CUDA_TRY(cudaStreamWaitEvent(ln.st, h->ev_fork, 0));
      }
      for (size_t k = 0; k < todo.size(); ++k) {
        auto& ln = h->lanes[k % n_lanes];
        CachedCloud& c = *todo[k];
        if (do_align && (!c.has_ndt || c.ndt_res != res)) B200_CUDA_TRY(cache_build_ndt(ln.st, ln.grid, c, res));
        if (with_fitness && !c.has_nn) B200_CUDA_TRY(cache_build_nn(ln.st, ln.nn, c));
      }
      for (int l = 0; l < n_lanes; ++l) {  // join
        B200_CUDA_TRY(cudaEventRecord(h->lanes[l].done, h->lanes[l].st));
        B200_CUDA_TRY(cudaStreamWaitEvent(h->stream, h->lanes[l].done, 0));
      }
    }
  }
  // ---- jobs (pairs with an empty source never reach the kernel: PCL's initCompute fails, converged_ stays false)
  B200_CUDA_TRY(h->pin_batch.reserve(n_pairs * (sizeof(NdtJob) + sizeof(FitJob) + sizeof(b200reg_result))));
  NdtJob* hj = reinterpret_cast<NdtJob*>(h->pin_batch.p);
  FitJob* hf = reinterpret_cast<FitJob*>(h->pin_batch.p + n_pairs * sizeof(NdtJob));
  b200reg_result* hr = reinterpret_cast<b200reg_result*>(h->pin_batch.p + n_pairs * (sizeof(NdtJob) + sizeof(FitJob)));
  B200_CUDA_TRY(h->jobs.reserve(n_pairs));
  B200_CUDA_TRY(h->batch_results.reserve(n_pairs));
  int n_jobs = 0;
  std::vector<int> job_pair;
  job_pair.reserve(n_pairs);
  for (size_t i = 0; i < n_pairs; ++i) {
    b200reg_result& r = hr[i];
    memset(&r, 0, sizeof(r));
    memcpy(r.transformation, pairs[i].guess, 64);
    r.fitness = 1.7976931348623157e308;
    if (src[i]->n == 0) continue;
    NdtJob& j = hj[n_jobs];
    memset(&j, 0, sizeof(j));
    j.src = src[i]->pts.p;
    j.n_src = src[i]->n;
    j.grid = tgt[i]->ndt_view();
    j.result = h->batch_results.p + i;
    const float* g = pairs[i].guess;
    for (int rr = 0; rr < 3; ++rr)
      for (int c = 0; c < 4; ++c) j.guess[4 * rr + c] = g[4 * c + rr];
    float eul[3];
    euler_xyz_from_colmajor(g, eul);
    j.p0[0] = g[12]; j.p0[1] = g[13]; j.p0[2] = g[14];
    j.p0[3] = eul[0]; j.p0[4] = eul[1]; j.p0[5] = eul[2];
    job_pair.push_back((int)i);
    ++n_jobs;
  }
  B200_CUDA_TRY(cudaMemcpyAsync(h->batch_results.p, hr, n_pairs * sizeof(b200reg_result), cudaMemcpyHostToDevice, h->stream));
  h->batch_align_ms = h->batch_fitness_ms = 0.0;
  if (n_jobs && do_align) {
    B200_CUDA_TRY(cudaMemcpyAsync(h->jobs.p, hj, (size_t)n_jobs * sizeof(NdtJob), cudaMemcpyHostToDevice, h->stream));
    // few pairs: several SMs cooperate on each; a full batch: one SM per registration, no grid-wide sync at all
    int G = h->num_sm / n_jobs;
    if (G < 1) {
      // more pairs than SMs: the batch runs in rounds of num_sm / G registrations, and the last round is
      // rarely full — 512 pairs on 148 SMs are 3.46 rounds of one-CTA registrations, i.e. 4.  Two or
      // four CTAs per registration halve / quarter a registration's time and waste less of the last round
      // (512 pairs: 7 rounds of half-length registrations = 3.5).  A group pays a barrier per pass (~1 %).
      double best_cost = 0.0;
      for (int g = 1; g <= 4; ++g) {
        const int groups = h->num_sm / g;
        const int rounds = (n_jobs + groups - 1) / groups;
        const double cost = (double)rounds * (1.0 / g) * (g > 1 ? 1.02 : 1.0);
        if (g == 1 || cost < best_cost * 0.97) { best_cost = cost; G = g; }
      }
    }
    const int n_groups = h->num_sm / G;
    B200_CUDA_TRY(h->partials.reserve((size_t)n_groups * 2 * G * kAccStride));
    if ((rc = ensure_barriers(h, (size_t)(n_groups + 1) * 32))) return rc;
    if ((rc = drain_events(h))) return rc;
    if ((rc = begin_timed_launch(h))) return rc;
    // more pairs than SMs: CTAs work through target runs and steal at the end (see NdtTargetQueue)
    NdtTargetQueue tq{nullptr, nullptr, 0};
    if (G == 1 && n_jobs > n_groups) {
      std::vector<uint2> runs;
      for (int j = 0; j < n_jobs; ++j) {
        if (j > 0 && hj[j].grid.table == hj[j - 1].grid.table) runs.back().y += 1u;
        else runs.push_back(make_uint2((unsigned)j, 1u));
      }
      B200_CUDA_TRY(h->pin_runs.reserve(runs.size()));
      memcpy(h->pin_runs.p, runs.data(), runs.size() * sizeof(uint2));
      B200_CUDA_TRY(h->tq_runs.reserve(runs.size()));
      B200_CUDA_TRY(h->tq_next.reserve(runs.size()));
      B200_CUDA_TRY(cudaMemcpyAsync(h->tq_runs.p, h->pin_runs.p, runs.size() * sizeof(uint2), cudaMemcpyHostToDevice, h->stream));
      B200_CUDA_TRY(cudaMemsetAsync(h->tq_next.p, 0, runs.size() * sizeof(unsigned int), h->stream));
      tq.runs = h->tq_runs.p; tq.next = h->tq_next.p; tq.n_runs = (int)runs.size();
    }
    B200_CUDA_TRY(launch_ndt_mode(h, n_jobs, G, n_groups, nullptr, tq));
    launch_counter() += 1;
    if ((rc = end_timed_launch(h))) return rc;
  }
  // ---- getFitnessScore(max_range) for every pair, in chunks that bound the d2 scratch
  cudaEvent_t evf0 = nullptr, evf1 = nullptr;
  if (with_fitness && n_jobs) {
    const float max_d2 = fitness_max_range >= 3.0e38 ? 3.402823466e+38f : (float)fitness_max_range * 1.0001f + 1e-6f;
    const long long kChunkPoints = 32ll << 20;
    B200_CUDA_TRY(h->fit_jobs.reserve(n_jobs));
    B200_CUDA_TRY(h->batch_n_pending.reserve(2));
    if (h->timing) { B200_CUDA_TRY(cudaEventCreate(&evf0)); B200_CUDA_TRY(cudaEventCreate(&evf1)); B200_CUDA_TRY(cudaEventRecord(evf0, h->stream)); }
    int j0 = 0;
    while (j0 < n_jobs) {
      long long pts = 0;
      int j1 = j0, max_n = 0;
      while (j1 < n_jobs && j1 - j0 < 65535 && (j1 == j0 || pts + src[job_pair[j1]]->n <= kChunkPoints)) {
        const int i = job_pair[j1];
        FitJob& f = hf[j1];
        f.view = tgt[i]->nn_view();
        f.src = src[i]->pts.p;
        f.n_src = src[i]->n;
        f.result = i;
        f.d2_offset = pts;
        pts += src[i]->n;
        if (src[i]->n > max_n) max_n = src[i]->n;
        ++j1;
      }
      const int nj = j1 - j0;
      B200_CUDA_TRY(h->batch_d2.reserve((size_t)pts));
      B200_CUDA_TRY(h->batch_pending.reserve((size_t)pts));
      B200_CUDA_TRY(h->batch_pending2.reserve((size_t)pts));
      B200_CUDA_TRY(cudaMemcpyAsync(h->fit_jobs.p + j0, hf + j0, (size_t)nj * sizeof(FitJob), cudaMemcpyHostToDevice, h->stream));
      B200_CUDA_TRY(cudaMemsetAsync(h->batch_n_pending.p, 0, 2 * sizeof(unsigned int), h->stream));
      launch_counter() += 4;
      k_nn_search_batch<<<dim3((max_n + 256 / B200_FIT_LANES - 1) / (256 / B200_FIT_LANES), nj), 256, 0, h->stream>>>(h->fit_jobs.p + j0, h->batch_results.p, max_d2, h->batch_d2.p, h->batch_pending.p, h->batch_n_pending.p);
      k_nn_far_batch<<<kNumSM * 8, 256, 0, h->stream>>>(h->fit_jobs.p + j0, h->batch_results.p, h->batch_pending.p, h->batch_n_pending.p, max_d2, h->batch_d2.p, h->batch_pending2.p,
                                                         h->batch_n_pending.p + 1);
      k_nn_bruteforce_batch<<<kNumSM * 8, 256, 0, h->stream>>>(h->fit_jobs.p + j0, h->batch_results.p, h->batch_pending2.p, h->batch_n_pending.p + 1, h->batch_d2.p);
      k_fitness_batch<<<nj, 256, 0, h->stream>>>(h->fit_jobs.p + j0, h->batch_d2.p, fitness_max_range, h->batch_results.p);
      B200_CUDA_TRY(cudaGetLastError());
      j0 = j1;
    }
    if (h->timing) B200_CUDA_TRY(cudaEventRecord(evf1, h->stream));
  }
  if (keep_on_device) {
    if (evf0) { cudaEventDestroy(evf0); cudaEventDestroy(evf1); }
    return B200REG_OK;
  }
  B200_CUDA_TRY(cudaMemcpyAsync(hr, h->batch_results.p, n_pairs * sizeof(b200reg_result), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA_TRY(cudaStreamSynchronize(h->stream));
  memcpy(results, hr, n_pairs * sizeof(b200reg_result));
  if (!with_fitness)
    for (size_t i = 0; i < n_pairs; ++i) results[i].fitness = 1.7976931348623157e308;
  if (h->timing && n_jobs && do_align) {
    const double before = h->align_ms;
    if ((rc = drain_events(h))) return rc;
    h->batch_align_ms = h->align_ms - before;
    if (evf0) {
      float ms = 0.f;
      B200_CUDA_TRY(cudaEventElapsedTime(&ms, evf0, evf1));
      h->batch_fitness_ms = (double)ms;
    }
  }
  if (evf0) { cudaEventDestroy(evf0); cudaEventDestroy(evf1); }
  return B200REG_OK;
}

// ---- cached clouds as the source / target of a plain align (the serial loop of a FAST_GICP LoopDetector) -----------------
// A candidate keyframe is registered against many new keyframes over a run [REF include/hdl_graph_slam/loop_detector.hpp:
// 83-111,124-156]; with its cloud in the cache (b200reg_cloud_put) a FAST_GICP handle computes its covariances — NN lattice,
// k-NN, regularisation: ~0.4 ms per 48 k-point cloud, more than the upload — once in the keyframe's life instead of once
// per pair.  Same kernels on the same cloud: the registration is bit-identical to b200reg_set_source(cloud) + align.
static int cached_covariances(b200reg_handle* h, CachedCloud& c) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  const int k = h->cfg.correspondence_randomness, reg = h->cfg.regularization;
  if (c.has_cov && c.cov_k == k && c.cov_reg == reg) return B200REG_OK;
  if (k > 32) { h->err = "reg_correspondence_randomness above 32 is not supported by the device k-NN (one neighbour per warp lane)"; return B200REG_E_INVALID; }
  if (!c.has_nn) B200_CUDA_TRY(cache_build_nn(h->stream, h->nn_src, c));  // nn_src serves as the builder: its own view goes stale
  h->nn_src_stale = true;
  B200_CUDA_TRY(c.cov.reserve((size_t)(c.n > 0 ? c.n : 1) * 6));
  B200_CUDA_TRY(launch_gicp_covariances(h, c.nn_view(), c.pts.p, c.n, c.cov.p));
  c.has_cov = true; c.cov_k = k; c.cov_reg = reg;
  return B200REG_OK;
}

int b200reg_set_source_cached(b200reg_handle* h, int64_t id) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h) return B200REG_E_INVALID;
  CachedCloud* c = cache_find(h, (long long)id);
  if (!c) { h->err = "set_source_cached: cloud id " + std::to_string((long long)id) + " was never put"; return B200REG_E_INVALID; }
  int rc = b200reg_set_source_device(h, reinterpret_cast<const float*>(c->pts.p), (size_t)c->n);
  if (rc) return rc;
  if (h->cfg.method != B200REG_METHOD_GICP || c->n == 0) return B200REG_OK;
  if ((rc = cached_covariances(h, *c))) return rc;
  B200_CUDA_TRY(h->cov_src.reserve((size_t)c->n * 6));
  B200_CUDA_TRY(cudaMemcpyAsync(h->cov_src.p, c->cov.p, (size_t)c->n * 48, cudaMemcpyDeviceToDevice, h->stream));
  h->cov_src_ok = true;
  return B200REG_OK;
}

int b200reg_set_target_cached(b200reg_handle* h, int64_t id) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h) return B200REG_E_INVALID;
  CachedCloud* c = cache_find(h, (long long)id);
  if (!c) { h->err = "set_target_cached: cloud id " + std::to_string((long long)id) + " was never put"; return B200REG_E_INVALID; }
  int rc = b200reg_set_target_device(h, reinterpret_cast<const float*>(c->pts.p), (size_t)c->n);
  if (rc) return rc;
  if (h->cfg.method != B200REG_METHOD_GICP || c->n == 0) return B200REG_OK;
  if ((rc = cached_covariances(h, *c))) return rc;
  B200_CUDA_TRY(h->cov_tgt.reserve((size_t)c->n * 6));
  B200_CUDA_TRY(cudaMemcpyAsync(h->cov_tgt.p, c->cov.p, (size_t)c->n * 48, cudaMemcpyDeviceToDevice, h->stream));
  h->cov_tgt_ok = true;
  return B200REG_OK;
}

int b200reg_align_batch(b200reg_handle* h, const b200reg_pair* pairs, size_t n_pairs, int with_fitness, double fitness_max_range, b200reg_result* results) {
  return batch_run(h, pairs, n_pairs, 1, with_fitness, fitness_max_range, results);
}

// internal (not in include/b200reg.h): the batch enqueued only, records left on the device — see b200reg_multi.cu
int b200reg_internal_align_batch_device(b200reg_handle* h, const b200reg_pair* pairs, size_t n_pairs, int with_fitness, double fitness_max_range, size_t min_slots, b200reg_result** d_results) {
  if (!d_results) return B200REG_E_INVALID;
  if (!n_pairs) {  // a device without a share still takes part in the gather: it needs a send buffer
    b200reg_pair none;
    (void)none;
    auto set_error = [&](const std::string& s) { h->err = s; };
    if (!h) return B200REG_E_INVALID;
    int rc = set_device(h);
    if (rc) return rc;
    B200_CUDA_TRY(h->batch_results.reserve(min_slots + 1));
    *d_results = h->batch_results.p;
    return B200REG_OK;
  }
  return batch_run(h, pairs, n_pairs, 1, with_fitness, fitness_max_range, nullptr, d_results, min_slots);
}

int b200reg_calc_fitness_batch(b200reg_handle* h, const b200reg_pair* pairs, size_t n_pairs, double max_range, double* out) {
  if (!h || (n_pairs && (!pairs || !out))) return B200REG_E_INVALID;
  std::vector<b200reg_result> res(n_pairs);
  const int rc = batch_run(h, pairs, n_pairs, 0, 1, max_range, res.data());
  if (rc) return rc;
  for (size_t i = 0; i < n_pairs; ++i) out[i] = res[i].fitness;
  return B200REG_OK;
}

int b200reg_get_batch_timing(b200reg_handle* h, double* align_kernel_ms, double* fitness_ms) {
  if (!h) return B200REG_E_INVALID;
  if (align_kernel_ms) *align_kernel_ms = h->batch_align_ms;
  if (fitness_ms) *fitness_ms = h->batch_fitness_ms;
  return B200REG_OK;
}

