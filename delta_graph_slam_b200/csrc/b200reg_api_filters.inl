// b200reg — the prefilter chain behind the C ABI: VoxelGrid (with the base_link transform and the distance gate in front),
// RadiusOutlierRemoval / StatisticalOutlierRemoval, the stand-alone distance filter and filtered2D
// [REF apps/prefiltering_nodelet.cpp:111-164, 198-291].  Included by b200reg_api.cu (one translation unit: the entry points
// share the handle definition and the file-local helpers above).
// ---- VoxelGrid -------------------------------------------------------------------------------
static int vg_run(b200reg_handle* h, const float4* d_in, size_t n, const float leaf[3], unsigned min_pts, int dense, float4* d_out, float4* host_out = nullptr, size_t host_cap = 0) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  B200_CUDA_TRY(h->vg_id.reserve(n ? n : 1));
  B200_CUDA_TRY(h->vg_count.reserve(n ? n : 1));
  B200_CUDA_TRY(h->vg_counts.reserve(1));
  B200_CUDA_TRY(h->vg_sorted.reserve(n ? n : 1));
  if (h->in_xf_on && n) {  // base_link frame first, as cloud_callback does [REF apps/prefiltering_nodelet.cpp:123-148]
    B200_CUDA_TRY(h->in_xf_buf.reserve(n));
    launch_counter() += 1;
    k_input_transform<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(d_in, (int)n, h->in_xf, dense, h->in_xf_buf.p);
    d_in = h->in_xf_buf.p;
  }
  bool gathered = false;  // the one-sweep sort's last pass writes the points in sorted order itself
  B200_CUDA_TRY(h->vg_sort.run(h->stream, d_in, (int)n, dense, leaf[0], leaf[1], leaf[2], true, h->gate, h->vg_sorted.p, &gathered));
  const int blocks = n ? (int)((n + 255) / 256) : 1;
  launch_counter() += (gathered ? 1 : 2) + (min_pts > 1 ? 1 : 0);
  if (!gathered) k_vg_gather<<<blocks, 256, 0, h->stream>>>(d_in, (int)n, h->vg_sort.vals_a.p, h->vg_sort.vals_b.p, h->vg_sort.meta.p, h->vg_sorted.p);
  // the overflow case publishes from k_vg_centroids even when a compaction pass follows
  VgCounts* hc = const_cast<VgCounts*>(&h->mail->vg);
  unsigned int* hf = const_cast<unsigned int*>(&h->mail->vg_seq);
  const unsigned int seq = ++h->vg_seq;
  if (!h->vg_done.p) {  // completion counter of k_vg_centroids: zeroed once, the last block restores it
    B200_CUDA_TRY(h->vg_done.reserve(1));
    B200_CUDA_TRY(cudaMemsetAsync(h->vg_done.p, 0, h->vg_done.cap * sizeof(unsigned int), h->stream));
  }
  k_vg_centroids<<<blocks, 256, 0, h->stream>>>(d_in, (int)n, h->vg_sort.vals_a.p, h->vg_sort.vals_b.p, h->vg_sorted.p, h->vg_sort.seg_first.p, h->vg_sort.meta.p, h->vg_sort.vox_start.p, h->vg_sort.vox_key.p,
                                                min_pts, d_out, h->vg_id.p, h->vg_count.p, h->vg_counts.p, hc, hf, seq, h->vg_done.p, min_pts > 1 ? 0 : 1, host_out,
                                                (unsigned)(host_cap > 0xFFFFFFFFull ? 0xFFFFFFFFull : host_cap), h->gate.on);
  if (h->gate.on) {  // only acts in the "leaf size too small" case: the output is then the gated input
    launch_counter() += 1;
    k_gate_copy<<<1, 1024, 0, h->stream>>>(d_in, (int)n, h->gate, h->vg_sort.meta.p, d_out, host_out, (unsigned)(host_cap > 0xFFFFFFFFull ? 0xFFFFFFFFull : host_cap), h->vg_counts.p, hc, hf, seq);
  }
  if (min_pts > 1) k_vg_compact<<<1, 1024, 0, h->stream>>>(h->vg_sort.meta.p, min_pts, d_out, h->vg_id.p, h->vg_count.p, h->vg_counts.p, hc, hf, seq);
  B200_CUDA_TRY(cudaGetLastError());
  h->vg_last_n = (int)n;
  return B200REG_OK;
}

// The filter as two halves.  begin: everything is enqueued on the handle's stream and the call
// returns; end: wait for the point count the last kernel publishes through the mailbox.  The
// synchronous entry points below are begin + end.
int b200reg_voxelgrid_filter_device_begin(b200reg_handle* h, const float* d_xyzw, size_t n, const float leaf[3], unsigned min_pts, int dense, float* d_out) {
  if (!h || !leaf || (n && (!d_xyzw || !d_out))) return B200REG_E_INVALID;
  if (!(leaf[0] > 0 && leaf[1] > 0 && leaf[2] > 0)) return B200REG_E_INVALID;
  if (h->vg_pending.active) { h->err = "a filter call is already in flight on this handle (b200reg_voxelgrid_filter_end first)"; return B200REG_E_STATE; }
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = vg_run(h, (const float4*)d_xyzw, n, leaf, min_pts, dense, (float4*)d_out))) return rc;
  h->vg_pending = b200reg_handle::VgPending();
  h->vg_pending.active = true;
  return B200REG_OK;
}

int b200reg_voxelgrid_filter_begin(b200reg_handle* h, const float* xyzw, size_t n, size_t stride, const float leaf[3], unsigned min_pts, int dense, float* out, size_t cap) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h || !leaf || (n && !xyzw)) return B200REG_E_INVALID;
  if (!(leaf[0] > 0 && leaf[1] > 0 && leaf[2] > 0)) return B200REG_E_INVALID;
  if (h->vg_pending.active) { h->err = "a filter call is already in flight on this handle (b200reg_voxelgrid_filter_end first)"; return B200REG_E_STATE; }
  if (stride < 12 || (stride % 4) != 0) { h->err = "stride_bytes must be a multiple of 4 and at least 12"; return B200REG_E_INVALID; }
  int rc = set_device(h);
  if (rc) return rc;
  // input: page-locked caller memory is read by DMA while this call has already returned (the caller
  // keeps it unchanged until _end); pageable memory goes through the handle's pinned staging buffer,
  // which is not touched again before the next _begin
  B200_CUDA_TRY(h->stage_in.reserve(n ? n : 1));
  B200_CUDA_TRY(h->stage_out.reserve(n ? n : 1));
  if (n) {
    if (stride == 16 && is_pinned_host(xyzw)) {
      B200_CUDA_TRY(cudaMemcpyAsync(h->stage_in.p, xyzw, n * 16, cudaMemcpyHostToDevice, h->stream));
    } else {
      B200_CUDA_TRY(h->pin_in.reserve(n));
      if (stride == 16) {
        host_copy(h->pin_in.p, xyzw, n * 16);
      } else {
        const unsigned char* b = (const unsigned char*)xyzw;
        for (size_t i = 0; i < n; ++i) {
          const float* p = (const float*)(b + i * stride);
          h->pin_in.p[i] = make_float4(p[0], p[1], p[2], 1.0f);
        }
      }
      B200_CUDA_TRY(cudaMemcpyAsync(h->stage_in.p, h->pin_in.p, n * 16, cudaMemcpyHostToDevice, h->stream));
    }
  }
  // output: a page-locked caller cloud is written by the centroid kernel itself (mapped memory);
  // min_points_per_voxel > 1 compacts on the device first, and pageable memory needs staging: both
  // copy in _end
  const bool zero_copy = out && cap && min_pts <= 1 && is_pinned_host(out);
  if ((rc = vg_run(h, h->stage_in.p, n, leaf, min_pts, dense, h->stage_out.p, zero_copy ? (float4*)out : nullptr, cap))) return rc;
  h->vg_pending.active = true;
  h->vg_pending.host_out = out;
  h->vg_pending.cap = cap;
  h->vg_pending.zero_copy = zero_copy;
  return B200REG_OK;
}

// host scan in, filtered cloud left on the device (a fused front end hands it to the registration without a round trip
// through host memory)
int b200reg_voxelgrid_filter_host_to_device_begin(b200reg_handle* h, const float* xyzw, size_t n, size_t stride, const float leaf[3], unsigned min_pts, int dense, float* d_out) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h || !leaf || (n && (!xyzw || !d_out))) return B200REG_E_INVALID;
  if (!(leaf[0] > 0 && leaf[1] > 0 && leaf[2] > 0)) return B200REG_E_INVALID;
  if (h->vg_pending.active) { h->err = "a filter call is already in flight on this handle (b200reg_voxelgrid_filter_end first)"; return B200REG_E_STATE; }
  if (stride < 12 || (stride % 4) != 0) { h->err = "stride_bytes must be a multiple of 4 and at least 12"; return B200REG_E_INVALID; }
  int rc = set_device(h);
  if (rc) return rc;
  B200_CUDA_TRY(h->stage_in.reserve(n ? n : 1));
  if (n) {
    if (stride == 16 && is_pinned_host(xyzw)) {
      B200_CUDA_TRY(cudaMemcpyAsync(h->stage_in.p, xyzw, n * 16, cudaMemcpyHostToDevice, h->stream));
    } else {
      B200_CUDA_TRY(h->pin_in.reserve(n));
      if (stride == 16) {
        host_copy(h->pin_in.p, xyzw, n * 16);
      } else {
        const unsigned char* b = (const unsigned char*)xyzw;
        for (size_t i = 0; i < n; ++i) {
          const float* p = (const float*)(b + i * stride);
          h->pin_in.p[i] = make_float4(p[0], p[1], p[2], 1.0f);
        }
      }
      B200_CUDA_TRY(cudaMemcpyAsync(h->stage_in.p, h->pin_in.p, n * 16, cudaMemcpyHostToDevice, h->stream));
    }
  }
  if ((rc = vg_run(h, h->stage_in.p, n, leaf, min_pts, dense, (float4*)d_out))) return rc;
  h->vg_pending = b200reg_handle::VgPending();
  h->vg_pending.active = true;
  return B200REG_OK;
}

int b200reg_voxelgrid_filter_end(b200reg_handle* h, size_t* n_out) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h || !n_out) return B200REG_E_INVALID;
  *n_out = 0;
  if (!h->vg_pending.active) { h->err = "no filter call in flight on this handle"; return B200REG_E_STATE; }
  const b200reg_handle::VgPending pend = h->vg_pending;
  h->vg_pending.active = false;
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = wait_mail(h, &h->mail->vg_seq, h->vg_seq))) return rc;
  const size_t m = h->mail->vg.n_out;
  *n_out = m;
  h->vg_last_out = (int)m;
  if (!pend.host_out && !pend.cap) return B200REG_OK;  // device variant
  if (m > pend.cap) { h->err = "output capacity too small"; return B200REG_E_CAPACITY; }
  if (!m || pend.zero_copy) return B200REG_OK;
  if (!pend.host_out) return B200REG_E_INVALID;
  if (is_pinned_host(pend.host_out)) {
    B200_CUDA_TRY(cudaMemcpyAsync(pend.host_out, h->stage_out.p, m * 16, cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA_TRY(cudaStreamSynchronize(h->stream));
  } else {
    B200_CUDA_TRY(h->pin_out.reserve(m));
    B200_CUDA_TRY(cudaMemcpyAsync(h->pin_out.p, h->stage_out.p, m * 16, cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA_TRY(cudaStreamSynchronize(h->stream));
    host_copy(pend.host_out, h->pin_out.p, m * 16);
  }
  return B200REG_OK;
}

int b200reg_voxelgrid_filter_device(b200reg_handle* h, const float* d_xyzw, size_t n, const float leaf[3], unsigned min_pts, int dense, float* d_out, size_t* n_out) {
  if (!n_out) return B200REG_E_INVALID;
  int rc = b200reg_voxelgrid_filter_device_begin(h, d_xyzw, n, leaf, min_pts, dense, d_out);
  if (rc) return rc;
  return b200reg_voxelgrid_filter_end(h, n_out);
}

int b200reg_voxelgrid_filter(b200reg_handle* h, const float* xyzw, size_t n, size_t stride, const float leaf[3], unsigned min_pts, int dense, float* out, size_t cap,
                             size_t* n_out) {
  if (!n_out) return B200REG_E_INVALID;
  *n_out = 0;
  int rc = b200reg_voxelgrid_filter_begin(h, xyzw, n, stride, leaf, min_pts, dense, out, cap);
  if (rc) return rc;
  return b200reg_voxelgrid_filter_end(h, n_out);
}

int b200reg_set_distance_filter(b200reg_handle* h, int use, double near_thresh, double far_thresh) {
  if (!h) return B200REG_E_INVALID;
  h->gate.on = use ? 1 : 0;
  h->gate.near_thresh = near_thresh;
  h->gate.far_thresh = far_thresh;
  return B200REG_OK;
}

int b200reg_set_input_transform(b200reg_handle* h, const double* matrix4x4_colmajor) {
  if (!h) return B200REG_E_INVALID;
  h->in_xf_on = matrix4x4_colmajor != nullptr;
  if (matrix4x4_colmajor)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) h->in_xf.m[4 * r + c] = matrix4x4_colmajor[4 * c + r];
  return B200REG_OK;
}

// ---- RadiusOutlierRemoval ----------------------------------------------------------------------
// which outlier filter a call runs: pcl::RadiusOutlierRemoval or pcl::StatisticalOutlierRemoval
struct OutlierSpec {
  bool statistical = false;
  double radius = 0.0;
  int min_neighbors = 0;
  int mean_k = 0;
  double stddev_mul = 0.0;
  // filtered2D of the prefilter nodelet: height gate -> normal test -> flatten (b200reg_flat_filter)
  bool flat = false;
  double lidar_z = 0.0;
  int normal_k = 0;
  float normal_thresh = 0.f;
  // distance_filter on its own (down-sampling NONE) [REF apps/prefiltering_nodelet.cpp:275-291]
  bool gate_only = false;
  double near_thresh = 0.0, far_thresh = 0.0;
  bool valid() const {
    if (gate_only) return near_thresh == near_thresh && far_thresh == far_thresh;
    if (flat) return normal_k >= 1 && normal_k <= 32 && lidar_z == lidar_z && normal_thresh == normal_thresh;
    return statistical ? (mean_k >= 1 && mean_k <= 31 && stddev_mul == stddev_mul) : (radius > 0 && min_neighbors >= 0);
  }
};

static int ror_run(b200reg_handle* h, const float4* d_in, size_t n, const OutlierSpec& spec, float4* d_out, float4* host_out, size_t host_cap) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  const double radius = spec.radius;
  const int min_neighbors = spec.min_neighbors;
  PointGate gate = kNoGate;
  if (spec.flat) { gate.on = 2; gate.near_thresh = spec.lidar_z; }  // height_filtering as a gate of the lattice build: no intermediate cloud
  if (spec.gate_only && h->in_xf_on && n) {  // distance_filter is the first stage of a prefilter without a down-sampler: base_link frame first
    B200_CUDA_TRY(h->in_xf_buf.reserve(n));
    launch_counter() += 1;
    k_input_transform<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(d_in, (int)n, h->in_xf, /*is_dense=*/0, h->in_xf_buf.p);
    d_in = h->in_xf_buf.p;
  }
  if (!spec.gate_only) B200_CUDA_TRY(h->nn_ror.build(h->stream, d_in, (int)n, /*is_dense=*/0, gate));
  const int blocks = n ? (int)((n + 255) / 256) : 1;
  B200_CUDA_TRY(h->ror_keep.reserve(n ? n : 1));
  B200_CUDA_TRY(h->ror_block_count.reserve(blocks));
  B200_CUDA_TRY(h->ror_counts.reserve(1));
  if (!h->ror_done.p) {
    B200_CUDA_TRY(h->ror_done.reserve(1));
    B200_CUDA_TRY(cudaMemsetAsync(h->ror_done.p, 0, h->ror_done.cap * sizeof(unsigned int), h->stream));
  }
  const float r2 = (float)(radius * radius);
  int rings = (int)ceil(radius / (double)kNnCell);
  if (rings < 1) rings = 1;
  RorCounts* hc = const_cast<RorCounts*>(&h->mail->ror);
  unsigned int* hf = const_cast<unsigned int*>(&h->mail->ror_seq);
  const unsigned int seq = ++h->ror_seq;
  if (spec.gate_only) {
    PointGate dg = {1, spec.near_thresh, spec.far_thresh};
    launch_counter() += 2;
    k_gate_flags<<<blocks, 256, 0, h->stream>>>(d_in, (int)n, dg, h->ror_keep.p, h->ror_block_count.p);
  } else if (spec.flat) {
    // |n_z| per point that passed the height gate; NaN (never kept) everywhere else
    B200_CUDA_TRY(h->sor_dist.reserve(n ? n : 1));
    B200_CUDA_TRY(h->sor_pending.reserve(n ? n : 1));
    B200_CUDA_TRY(h->sor_n_pending.reserve(1));
    B200_CUDA_TRY(cudaMemsetAsync(h->sor_dist.p, 0xFF, (n ? n : 1) * sizeof(float), h->stream));
    B200_CUDA_TRY(cudaMemsetAsync(h->sor_n_pending.p, 0, sizeof(unsigned int), h->stream));
    launch_counter() += 4;
    if (n) {
      k_gicp_knn<kKnnNormalNz><<<(int)((n + 7) / 8), 256, 0, h->stream>>>(h->nn_ror.view(), d_in, (int)n, spec.normal_k, nullptr, h->sor_pending.p, h->sor_n_pending.p, h->sor_dist.p);
      k_gicp_knn_brute<kKnnNormalNz><<<kNumSM * 2, kBruteWarps * 32, 0, h->stream>>>(h->nn_ror.view(), d_in, spec.normal_k, nullptr, h->sor_pending.p, h->sor_n_pending.p, h->sor_dist.p);
    }
    k_nz_flags<<<blocks, 256, 0, h->stream>>>(h->sor_dist.p, (int)n, spec.normal_thresh, h->ror_keep.p, h->ror_block_count.p);
  } else if (spec.statistical) {
    // the k-NN kernels write one float per finite point; everything else stays at the "not counted" mark (< 0)
    B200_CUDA_TRY(h->sor_dist.reserve(n ? n : 1));
    B200_CUDA_TRY(h->sor_stats.reserve(1));
    B200_CUDA_TRY(h->sor_pending.reserve(n ? n : 1));
    B200_CUDA_TRY(h->sor_n_pending.reserve(1));
    B200_CUDA_TRY(cudaMemsetAsync(h->sor_dist.p, 0xBF, (n ? n : 1) * sizeof(float), h->stream));
    B200_CUDA_TRY(cudaMemsetAsync(h->sor_n_pending.p, 0, sizeof(unsigned int), h->stream));
    launch_counter() += 5;
    if (n) {
      k_gicp_knn<kKnnMeanDistance><<<(int)((n + 7) / 8), 256, 0, h->stream>>>(h->nn_ror.view(), d_in, (int)n, spec.mean_k + 1, nullptr, h->sor_pending.p, h->sor_n_pending.p, h->sor_dist.p);
      k_gicp_knn_brute<kKnnMeanDistance><<<kNumSM * 2, kBruteWarps * 32, 0, h->stream>>>(h->nn_ror.view(), d_in, spec.mean_k + 1, nullptr, h->sor_pending.p, h->sor_n_pending.p, h->sor_dist.p);
    }
    k_sor_threshold<<<1, 1024, 0, h->stream>>>(h->sor_dist.p, (int)n, spec.stddev_mul, h->sor_stats.p);
    k_sor_flags<<<blocks, 256, 0, h->stream>>>(h->sor_dist.p, (int)n, h->sor_stats.p, h->ror_keep.p, h->ror_block_count.p);
  } else {
    launch_counter() += 2;
    k_ror_flags<<<blocks, 256, 0, h->stream>>>(h->nn_ror.view(), d_in, (int)n, r2, rings, min_neighbors, h->ror_keep.p, h->ror_block_count.p);
  }
  k_ror_scatter<<<blocks, 256, 0, h->stream>>>(d_in, (int)n, h->ror_keep.p, h->ror_block_count.p, d_out, host_out, (unsigned)(host_cap > 0xFFFFFFFFull ? 0xFFFFFFFFull : host_cap),
                                               h->ror_counts.p, hc, hf, seq, h->ror_done.p, spec.gate_only ? nullptr : h->nn_ror.sort.meta.p, spec.flat ? 1 : 0);
  B200_CUDA_TRY(cudaGetLastError());
  return B200REG_OK;
}

static int outlier_device_begin(b200reg_handle* h, const float* d_xyzw, size_t n, const OutlierSpec& spec, float* d_out) {
  if (!h || !spec.valid() || (n && (!d_xyzw || !d_out))) return B200REG_E_INVALID;
  if (h->ror_pending.active) { h->err = "an outlier-removal call is already in flight on this handle"; return B200REG_E_STATE; }
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = ror_run(h, (const float4*)d_xyzw, n, spec, (float4*)d_out, nullptr, 0))) return rc;
  h->ror_pending = b200reg_handle::RorPending();
  h->ror_pending.active = true;
  h->ror_pending.device = true;
  return B200REG_OK;
}

static OutlierSpec radius_spec(double radius, int min_neighbors) {
  OutlierSpec s;
  s.radius = radius; s.min_neighbors = min_neighbors;
  return s;
}
static OutlierSpec statistical_spec(int mean_k, double stddev_mul) {
  OutlierSpec s;
  s.statistical = true; s.mean_k = mean_k; s.stddev_mul = stddev_mul;
  return s;
}
int b200reg_radius_outlier_removal_device_begin(b200reg_handle* h, const float* d_xyzw, size_t n, double radius, int min_neighbors, float* d_out) {
  return outlier_device_begin(h, d_xyzw, n, radius_spec(radius, min_neighbors), d_out);
}
int b200reg_statistical_outlier_removal_device_begin(b200reg_handle* h, const float* d_xyzw, size_t n, int mean_k, double stddev_mul, float* d_out) {
  if (h && (mean_k < 1 || mean_k > 31)) { h->err = "statistical_mean_k must lie in 1..31 (the device k-NN holds one neighbour per warp lane)"; return B200REG_E_INVALID; }
  return outlier_device_begin(h, d_xyzw, n, statistical_spec(mean_k, stddev_mul), d_out);
}

static int outlier_host_begin(b200reg_handle* h, const float* xyzw, size_t n, size_t stride, const OutlierSpec& spec, float* out, size_t cap) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h || !spec.valid() || (n && !xyzw)) return B200REG_E_INVALID;
  if (h->ror_pending.active) { h->err = "an outlier-removal call is already in flight on this handle"; return B200REG_E_STATE; }
  if (stride < 12 || (stride % 4) != 0) { h->err = "stride_bytes must be a multiple of 4 and at least 12"; return B200REG_E_INVALID; }
  int rc = set_device(h);
  if (rc) return rc;
  B200_CUDA_TRY(h->ror_in.reserve(n ? n : 1));
  B200_CUDA_TRY(h->ror_out.reserve(n ? n : 1));
  if (n) {
    if (stride == 16 && is_pinned_host(xyzw)) {
      B200_CUDA_TRY(cudaMemcpyAsync(h->ror_in.p, xyzw, n * 16, cudaMemcpyHostToDevice, h->stream));
    } else {
      B200_CUDA_TRY(h->ror_pin_in.reserve(n));
      const unsigned char* b = (const unsigned char*)xyzw;
      for (size_t i = 0; i < n; ++i) {
        const float* p = (const float*)(b + i * stride);
        h->ror_pin_in.p[i] = make_float4(p[0], p[1], p[2], stride >= 16 ? p[3] : 1.0f);
      }
      B200_CUDA_TRY(cudaMemcpyAsync(h->ror_in.p, h->ror_pin_in.p, n * 16, cudaMemcpyHostToDevice, h->stream));
    }
  }
  const bool zero_copy = out && cap && is_pinned_host(out);
  if ((rc = ror_run(h, h->ror_in.p, n, spec, h->ror_out.p, zero_copy ? (float4*)out : nullptr, cap))) return rc;
  h->ror_pending = b200reg_handle::RorPending();
  h->ror_pending.active = true;
  h->ror_pending.host_out = out;
  h->ror_pending.cap = cap;
  h->ror_pending.zero_copy = zero_copy;
  return B200REG_OK;
}

int b200reg_radius_outlier_removal_begin(b200reg_handle* h, const float* xyzw, size_t n, size_t stride, double radius, int min_neighbors, float* out, size_t cap) {
  return outlier_host_begin(h, xyzw, n, stride, radius_spec(radius, min_neighbors), out, cap);
}
int b200reg_statistical_outlier_removal_begin(b200reg_handle* h, const float* xyzw, size_t n, size_t stride, int mean_k, double stddev_mul, float* out, size_t cap) {
  if (h && (mean_k < 1 || mean_k > 31)) { h->err = "statistical_mean_k must lie in 1..31 (the device k-NN holds one neighbour per warp lane)"; return B200REG_E_INVALID; }
  return outlier_host_begin(h, xyzw, n, stride, statistical_spec(mean_k, stddev_mul), out, cap);
}

int b200reg_radius_outlier_removal_end(b200reg_handle* h, size_t* n_out) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h || !n_out) return B200REG_E_INVALID;
  *n_out = 0;
  if (!h->ror_pending.active) { h->err = "no outlier-removal call in flight on this handle"; return B200REG_E_STATE; }
  const b200reg_handle::RorPending pend = h->ror_pending;
  h->ror_pending.active = false;
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = wait_mail(h, &h->mail->ror_seq, h->ror_seq))) return rc;
  if (h->mail->ror.overflow) {
    h->err = "outlier removal: the cloud spans more than 2^31 cells of the 0.5 m search lattice (gate it with the distance filter first); no result";
    return B200REG_E_INVALID;
  }
  const size_t m = h->mail->ror.n_out;
  *n_out = m;
  if (pend.device) return B200REG_OK;
  if (m > pend.cap) { h->err = "output capacity too small"; return B200REG_E_CAPACITY; }
  if (!m || pend.zero_copy) return B200REG_OK;
  if (!pend.host_out) return B200REG_E_INVALID;
  B200_CUDA_TRY(h->ror_pin_out.reserve(m));
  B200_CUDA_TRY(cudaMemcpyAsync(h->ror_pin_out.p, h->ror_out.p, m * 16, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA_TRY(cudaStreamSynchronize(h->stream));
  host_copy(pend.host_out, h->ror_pin_out.p, m * 16);
  return B200REG_OK;
}

int b200reg_radius_outlier_removal(b200reg_handle* h, const float* xyzw, size_t n, size_t stride, double radius, int min_neighbors, float* out, size_t cap, size_t* n_out) {
  if (!n_out) return B200REG_E_INVALID;
  *n_out = 0;
  int rc = b200reg_radius_outlier_removal_begin(h, xyzw, n, stride, radius, min_neighbors, out, cap);
  if (rc) return rc;
  return b200reg_radius_outlier_removal_end(h, n_out);
}

int b200reg_radius_outlier_removal_device(b200reg_handle* h, const float* d_xyzw, size_t n, double radius, int min_neighbors, float* d_out, size_t* n_out) {
  if (!n_out) return B200REG_E_INVALID;
  *n_out = 0;
  int rc = b200reg_radius_outlier_removal_device_begin(h, d_xyzw, n, radius, min_neighbors, d_out);
  if (rc) return rc;
  return b200reg_radius_outlier_removal_end(h, n_out);
}

int b200reg_statistical_outlier_removal_end(b200reg_handle* h, size_t* n_out) { return b200reg_radius_outlier_removal_end(h, n_out); }

int b200reg_statistical_outlier_removal(b200reg_handle* h, const float* xyzw, size_t n, size_t stride, int mean_k, double stddev_mul, float* out, size_t cap, size_t* n_out) {
  if (!n_out) return B200REG_E_INVALID;
  *n_out = 0;
  int rc = b200reg_statistical_outlier_removal_begin(h, xyzw, n, stride, mean_k, stddev_mul, out, cap);
  if (rc) return rc;
  return b200reg_radius_outlier_removal_end(h, n_out);
}

int b200reg_statistical_outlier_removal_device(b200reg_handle* h, const float* d_xyzw, size_t n, int mean_k, double stddev_mul, float* d_out, size_t* n_out) {
  if (!n_out) return B200REG_E_INVALID;
  *n_out = 0;
  int rc = b200reg_statistical_outlier_removal_device_begin(h, d_xyzw, n, mean_k, stddev_mul, d_out);
  if (rc) return rc;
  return b200reg_radius_outlier_removal_end(h, n_out);
}

// ---- distance_filter on its own (the prefilter nodelet with downsample_method NONE) ------------------
static OutlierSpec gate_spec(double near_thresh, double far_thresh) {
  OutlierSpec s;
  s.gate_only = true; s.near_thresh = near_thresh; s.far_thresh = far_thresh;
  return s;
}
int b200reg_distance_filter(b200reg_handle* h, const float* xyzw, size_t n, size_t stride, double near_thresh, double far_thresh, float* out, size_t cap, size_t* n_out) {
  if (!n_out) return B200REG_E_INVALID;
  *n_out = 0;
  int rc = outlier_host_begin(h, xyzw, n, stride, gate_spec(near_thresh, far_thresh), out, cap);
  if (rc) return rc;
  return b200reg_radius_outlier_removal_end(h, n_out);
}
int b200reg_distance_filter_device(b200reg_handle* h, const float* d_xyzw, size_t n, double near_thresh, double far_thresh, float* d_out, size_t* n_out) {
  if (!n_out) return B200REG_E_INVALID;
  *n_out = 0;
  int rc = outlier_device_begin(h, d_xyzw, n, gate_spec(near_thresh, far_thresh), d_out);
  if (rc) return rc;
  return b200reg_radius_outlier_removal_end(h, n_out);
}

// ---- filtered2D: height_filtering -> normal_filtering -> flatten ----------------------------------
static OutlierSpec flat_spec(double lidar_z, int k, double thresh) {
  OutlierSpec s;
  s.flat = true; s.lidar_z = lidar_z; s.normal_k = k; s.normal_thresh = (float)thresh;
  return s;
}
static int flat_check(b200reg_handle* h, int k) {
  if (h && (k < 1 || k > 32)) { h->err = "normal_k must lie in 1..32 (the device k-NN holds one neighbour per warp lane)"; return B200REG_E_INVALID; }
  return B200REG_OK;
}
int b200reg_flat_filter_begin(b200reg_handle* h, const float* xyzw, size_t n, size_t stride, double lidar_z, int normal_k, double normal_thresh, float* out, size_t cap) {
  int rc = flat_check(h, normal_k);
  if (rc) return rc;
  return outlier_host_begin(h, xyzw, n, stride, flat_spec(lidar_z, normal_k, normal_thresh), out, cap);
}
int b200reg_flat_filter_device_begin(b200reg_handle* h, const float* d_xyzw, size_t n, double lidar_z, int normal_k, double normal_thresh, float* d_out) {
  int rc = flat_check(h, normal_k);
  if (rc) return rc;
  return outlier_device_begin(h, d_xyzw, n, flat_spec(lidar_z, normal_k, normal_thresh), d_out);
}
int b200reg_flat_filter_end(b200reg_handle* h, size_t* n_out) { return b200reg_radius_outlier_removal_end(h, n_out); }
int b200reg_flat_filter(b200reg_handle* h, const float* xyzw, size_t n, size_t stride, double lidar_z, int normal_k, double normal_thresh, float* out, size_t cap, size_t* n_out) {
  if (!n_out) return B200REG_E_INVALID;
  *n_out = 0;
  int rc = b200reg_flat_filter_begin(h, xyzw, n, stride, lidar_z, normal_k, normal_thresh, out, cap);
  if (rc) return rc;
  return b200reg_radius_outlier_removal_end(h, n_out);
}
int b200reg_flat_filter_device(b200reg_handle* h, const float* d_xyzw, size_t n, double lidar_z, int normal_k, double normal_thresh, float* d_out, size_t* n_out) {
  if (!n_out) return B200REG_E_INVALID;
  *n_out = 0;
  int rc = b200reg_flat_filter_device_begin(h, d_xyzw, n, lidar_z, normal_k, normal_thresh, d_out);
  if (rc) return rc;
  return b200reg_radius_outlier_removal_end(h, n_out);
}
// |n_z| per input point of the last flat-filter call (NaN: below the height gate, or no normal)
int b200reg_flat_filter_last_nz(b200reg_handle* h, float* nz, size_t n) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h || !nz) return B200REG_E_INVALID;
  if (!h->sor_dist.p || n > h->sor_dist.cap) return B200REG_E_STATE;
  int rc = set_device(h);
  if (rc) return rc;
  B200_CUDA_TRY(cudaMemcpyAsync(nz, h->sor_dist.p, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return B200REG_OK;
}

// mean / stddev / cut of the last statistical call (after its _end), its count of points with a full
// neighbour list and whether the index-order summation had to run; dist (optional, n floats): the per-point
// mean neighbour distances, 0 for the points upstream leaves uncounted
int b200reg_statistical_last_stats(b200reg_handle* h, double stats3[3], unsigned long long* valid, int* exact_pass, float* dist, size_t n) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h) return B200REG_E_INVALID;
  if (!h->sor_stats.p) return B200REG_E_STATE;
  int rc = set_device(h);
  if (rc) return rc;
  SorStats st;
  B200_CUDA_TRY(cudaMemcpyAsync(&st, h->sor_stats.p, sizeof(st), cudaMemcpyDeviceToHost, h->stream));
  if (dist && n) {
    if (n > h->sor_dist.cap) return B200REG_E_INVALID;
    B200_CUDA_TRY(cudaMemcpyAsync(dist, h->sor_dist.p, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  }
  B200_CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (dist) for (size_t i = 0; i < n; ++i) if (dist[i] < 0.f) dist[i] = 0.f;
  if (stats3) { stats3[0] = st.mean; stats3[1] = st.stddev; stats3[2] = st.threshold; }
  if (valid) *valid = st.valid;
  if (exact_pass) *exact_pass = (int)st.exact_pass;
  return B200REG_OK;
}

int b200reg_voxelgrid_last_layout(b200reg_handle* h, uint32_t* voxel_id, uint32_t* count, size_t n_vox, uint32_t* key, size_t n_points, int32_t* grid6, int* overflow) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h) return B200REG_E_INVALID;
  if (!h->vg_sort.meta.p) return B200REG_E_STATE;
  int rc = set_device(h);
  if (rc) return rc;
  SortMeta meta;
  B200_CUDA_TRY(cudaMemcpyAsync(&meta, h->vg_sort.meta.p, sizeof(meta), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (overflow) *overflow = meta.grid.overflow;
  if (grid6) for (int a = 0; a < 3; ++a) { grid6[a] = meta.grid.min_b[a]; grid6[3 + a] = meta.grid.div_b[a]; }
  if (meta.grid.overflow) return B200REG_OK;
  size_t nv = n_vox < (size_t)h->vg_last_out ? n_vox : (size_t)h->vg_last_out;
  if (voxel_id && nv) B200_CUDA_TRY(cudaMemcpyAsync(voxel_id, h->vg_id.p, nv * 4, cudaMemcpyDeviceToHost, h->stream));
  if (count && nv) B200_CUDA_TRY(cudaMemcpyAsync(count, h->vg_count.p, nv * 4, cudaMemcpyDeviceToHost, h->stream));
  size_t np = n_points < (size_t)h->vg_last_n ? n_points : (size_t)h->vg_last_n;
  if (key && np) B200_CUDA_TRY(cudaMemcpyAsync(key, h->vg_sort.point_key.p, np * 4, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return B200REG_OK;
}

