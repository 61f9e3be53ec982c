// b200reg_api_map.inl — C entry points of the map-cloud path (map_cloud.cuh); included by b200reg_api.cu inside its
// extern "C" block (the handle type lives in that translation unit).
//
// MapCloudGenerator::generate [REF src/hdl_graph_slam/map_cloud_generator.cpp:13-49], called on a map-save / map-publish
// request with every keyframe snapshot of the graph [REF apps/delta_graph_slam_nodelet.cpp map_points_publish_timer_callback,
// save_map_service].

namespace {

struct MapSource {
  const float4* d_pts;  // device
  size_t n;
};

// the sequential part of pcl's octree, on the host: grow the box for one point exactly as adoptBoundingBoxToPoint /
// getKeyBitSize do (the arithmetic of oracle_capi.cpp orc_map_cloud, which restates pcl 1.8-1.10)
struct MapGrowth {
  double mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
  bool defined = false;
  unsigned depth = 0;
  unsigned long long shift[3] = {0, 0, 0};
  void admit(const float p[3], double resolution) {
    const float min_value = 1.1920928955078125e-07f;  // std::numeric_limits<float>::epsilon()
    while (true) {
      bool up[3], any = false;
      for (int a = 0; a < 3; ++a) {
        const bool lo = p[a] < mn[a];
        up[a] = p[a] >= mx[a];
        any = any || lo || up[a];
      }
      if (!any && defined) return;
      if (defined) {
        double side = static_cast<double>(1ull << depth) * resolution;
        for (int a = 0; a < 3; ++a)
          if (!up[a]) { mn[a] -= side; shift[a] += 1ull << depth; }
        ++depth;
        side = static_cast<double>(1ull << depth) * resolution - min_value;
        for (int a = 0; a < 3; ++a) mx[a] = mn[a] + side;
      } else {
        for (int a = 0; a < 3; ++a) { mn[a] = p[a] - resolution / 2; mx[a] = p[a] + resolution / 2; }
        unsigned max_key = 2;
        for (int a = 0; a < 3; ++a) max_key = std::max(max_key, static_cast<unsigned>(std::ceil((mx[a] - mn[a] - min_value) / resolution)));
        depth = static_cast<unsigned>(std::ceil(std::log2(static_cast<double>(max_key)) - min_value));
        const double side = static_cast<double>(1ull << depth) * resolution;
        for (int a = 0; a < 3; ++a) {
          const double over = (side - (mx[a] - mn[a])) / 2.0;
          if (over > min_value) { mn[a] -= over; mx[a] += over; }
        }
        defined = true;
      }
      if (depth > 62) return;  // caught by the caller's depth check
    }
  }
};

int map_cloud_run(b200reg_handle* h, const std::vector<MapSource>& src, const float* poses16, double resolution, float* out_xyzw, size_t out_capacity, size_t* n_out, double* min3_depth) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  const size_t n_kf = src.size();
  long long n = 0;
  std::vector<MapKeyframe> kfs(n_kf);
  for (size_t k = 0; k < n_kf; ++k) {
    kfs[k].src = src[k].d_pts;
    kfs[k].first = n;
    memcpy(kfs[k].T, poses16 + 16 * k, 64);
    n += (long long)src[k].n;
  }
  *n_out = 0;
  if (n > 0x7FFFFFF0ll) { h->err = "map cloud: more than 2^31 points"; return B200REG_E_INVALID; }
  if (n == 0) return B200REG_OK;
  // keyframes without points would break the "last keyframe whose first point <= i" search only if they came last with the
  // same `first` as a non-empty one; the search picks the LAST such entry, so empty keyframes are moved out of the table
  std::vector<MapKeyframe> table;
  for (size_t k = 0; k < n_kf; ++k)
    if (src[k].n) table.push_back(kfs[k]);
  B200_CUDA_TRY(h->map_kf.reserve(table.size()));
  B200_CUDA_TRY(h->map_world.reserve((size_t)n));
  B200_CUDA_TRY(cudaMemcpyAsync(h->map_kf.p, table.data(), table.size() * sizeof(MapKeyframe), cudaMemcpyHostToDevice, h->stream));
  const int blocks = (int)((n + 255) / 256);
  launch_counter() += 1;
  k_map_transform<<<blocks, 256, 0, h->stream>>>(h->map_kf.p, (int)table.size(), n, h->map_world.p);
  B200_CUDA_TRY(cudaGetLastError());
  if (!(resolution > 0.0)) {  // "to get unfiltered point cloud" (:33-34)
    if ((size_t)n > out_capacity) { h->err = "output capacity too small"; return B200REG_E_CAPACITY; }
    B200_CUDA_TRY(cudaMemcpyAsync(out_xyzw, h->map_world.p, (size_t)n * 16, cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA_TRY(cudaStreamSynchronize(h->stream));  // also keeps `table` alive until the upload has been consumed
    *n_out = (size_t)n;
    return B200REG_OK;
  }
  // ---- the box: one round per growth event
  B200_CUDA_TRY(h->map_scalar.reserve(4));
  MapGrowth g;
  MapEvents ev;
  memset(&ev, 0, sizeof(ev));
  long long start = 0;
  while (true) {
    const unsigned long long init = (unsigned long long)n;
    B200_CUDA_TRY(cudaMemcpyAsync(h->map_scalar.p, &init, 8, cudaMemcpyHostToDevice, h->stream));
    MapBox box;
    for (int a = 0; a < 3; ++a) { box.mn[a] = g.mn[a]; box.mx[a] = g.mx[a]; }
    box.defined = g.defined ? 1 : 0;
    long long span = n - start;
    int fb = (int)std::min<long long>((span + 255) / 256, (long long)kNumSM * 8);
    if (fb < 1) fb = 1;
    launch_counter() += 1;
    k_map_first_outside<<<fb, 256, 0, h->stream>>>(h->map_world.p, start, n, box, h->map_scalar.p);
    unsigned long long found = 0;
    B200_CUDA_TRY(cudaMemcpyAsync(&found, h->map_scalar.p, 8, cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (found >= (unsigned long long)n) break;
    float4 p;
    B200_CUDA_TRY(cudaMemcpy(&p, h->map_world.p + found, 16, cudaMemcpyDeviceToHost));
    const float pc[3] = {p.x, p.y, p.z};
    g.admit(pc, resolution);
    if (g.depth > (unsigned)kMapMaxDepth) { h->err = "map cloud: the map spans more than 2^21 voxels per axis at this resolution"; return B200REG_E_INVALID; }
    if (ev.n >= kMapMaxEvents) { h->err = "map cloud: too many bounding-box growth events"; return B200REG_E_INVALID; }
    ev.first[ev.n] = (long long)found;
    for (int a = 0; a < 3; ++a) { ev.mn[ev.n][a] = g.mn[a]; ev.shift[ev.n][a] = g.shift[a]; }
    ev.n += 1;
    start = (long long)found + 1;
  }
  if (min3_depth) { min3_depth[0] = g.mn[0]; min3_depth[1] = g.mn[1]; min3_depth[2] = g.mn[2]; min3_depth[3] = (double)g.depth; }
  if (ev.n == 0) return B200REG_OK;  // no finite point
  ev.first[ev.n] = n;
  for (int a = 0; a < 3; ++a) ev.shift_final[a] = g.shift[a];
  ev.resolution = resolution;
  ev.depth = (int)g.depth;
  // ---- keys -> sort -> distinct -> centres
  B200_CUDA_TRY(h->map_events.reserve(1));
  B200_CUDA_TRY(cudaMemcpyAsync(h->map_events.p, &ev, sizeof(ev), cudaMemcpyHostToDevice, h->stream));
  B200_CUDA_TRY(h->map_codes_a.reserve((size_t)n));
  B200_CUDA_TRY(h->map_codes_b.reserve((size_t)n));
  launch_counter() += 4;
  k_map_keys<<<blocks, 256, 0, h->stream>>>(h->map_world.p, n, h->map_events.p, h->map_codes_a.p);
  // the invalid code (all ones) must sort last: sort all 64 bits only when a non-finite point exists is not known here,
  // so the sort covers 3 * depth bits and one more (bit 63 of the invalid code); valid codes have it clear
  const uint32_t nbits = 64u;
  uint32_t* d_nbits = reinterpret_cast<uint32_t*>(h->map_scalar.p + 1);
  B200_CUDA_TRY(cudaMemcpyAsync(d_nbits, &nbits, 4, cudaMemcpyHostToDevice, h->stream));
  // valid codes use the low 3 * depth bits; digits above them are all zero for valid codes and all ones for the invalid
  // one, so a pass over such a digit only moves the invalid codes behind the valid ones — needed once, not per digit:
  // passes = digits holding code bits, plus the top digit
  const int code_passes = (3 * ev.depth + kOsRadixBits - 1) / kOsRadixBits;
  B200_CUDA_TRY(map_sort_codes(h, (int)n, code_passes));
  const bool in_b = ((code_passes + 1) & 1) != 0;
  const unsigned long long* sorted = in_b ? h->map_codes_b.p : h->map_codes_a.p;
  const long long n_tiles = (n + 1023) / 1024;
  B200_CUDA_TRY(h->map_tile_heads.reserve((size_t)n_tiles));
  k_map_heads<<<(int)n_tiles, 1024, 0, h->stream>>>(sorted, n, h->map_tile_heads.p);
  k_map_scan<<<1, 1024, 0, h->stream>>>(h->map_tile_heads.p, n_tiles, h->map_scalar.p + 2);
  B200_CUDA_TRY(h->map_out.reserve((size_t)n));
  k_map_centres<<<(int)n_tiles, 1024, 0, h->stream>>>(sorted, n, h->map_tile_heads.p, h->map_events.p, h->map_out.p, (unsigned long long)n);
  B200_CUDA_TRY(cudaGetLastError());
  unsigned long long m = 0;
  B200_CUDA_TRY(cudaMemcpyAsync(&m, h->map_scalar.p + 2, 8, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (m > out_capacity) { h->err = "output capacity too small"; *n_out = (size_t)m; return B200REG_E_CAPACITY; }
  if (m) B200_CUDA_TRY(cudaMemcpy(out_xyzw, h->map_out.p, (size_t)m * 16, cudaMemcpyDeviceToHost));
  *n_out = (size_t)m;
  return B200REG_OK;
}

}  // namespace

int b200reg_map_cloud(b200reg_handle* h, const float* const* clouds, const size_t* n_points, const float* poses16, size_t n_keyframes, double resolution, float* out_xyzw, size_t out_capacity,
                      size_t* n_out, double* min3_depth) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h || !n_out || (n_keyframes && (!clouds || !n_points || !poses16))) return B200REG_E_INVALID;
  *n_out = 0;
  if (!n_keyframes) { h->err = "warning: keyframes empty!!"; return B200REG_E_INVALID; }  // the reference prints this and returns nullptr (:14-17)
  int rc = set_device(h);
  if (rc) return rc;
  size_t total = 0;
  for (size_t k = 0; k < n_keyframes; ++k) {
    if (n_points[k] && !clouds[k]) return B200REG_E_INVALID;
    total += n_points[k];
  }
  if (total && !out_xyzw) return B200REG_E_INVALID;
  // the keyframe clouds go up into one staging area, one copy per keyframe (pageable memory is staged by the runtime)
  B200_CUDA_TRY(h->map_src.reserve(total ? total : 1));
  std::vector<MapSource> src(n_keyframes);
  size_t off = 0;
  for (size_t k = 0; k < n_keyframes; ++k) {
    if (n_points[k]) B200_CUDA_TRY(cudaMemcpyAsync(h->map_src.p + off, clouds[k], n_points[k] * 16, cudaMemcpyHostToDevice, h->stream));
    src[k] = MapSource{h->map_src.p + off, n_points[k]};
    off += n_points[k];
  }
  return map_cloud_run(h, src, poses16, resolution, out_xyzw, out_capacity, n_out, min3_depth);
}

int b200reg_map_cloud_cached(b200reg_handle* h, const int64_t* ids, const float* poses16, size_t n_keyframes, double resolution, float* out_xyzw, size_t out_capacity, size_t* n_out,
                             double* min3_depth) {
  if (!h || !n_out || (n_keyframes && (!ids || !poses16))) return B200REG_E_INVALID;
  *n_out = 0;
  if (!n_keyframes) { h->err = "warning: keyframes empty!!"; return B200REG_E_INVALID; }
  int rc = set_device(h);
  if (rc) return rc;
  std::vector<MapSource> src(n_keyframes);
  for (size_t k = 0; k < n_keyframes; ++k) {
    CachedCloud* c = cache_find(h, ids[k]);
    if (!c) { h->err = "map cloud: keyframe " + std::to_string(k) + " names a cloud id that was never put"; return B200REG_E_INVALID; }
    src[k] = MapSource{c->pts.p, (size_t)c->n};
  }
  return map_cloud_run(h, src, poses16, resolution, out_xyzw, out_capacity, n_out, min3_depth);
}
