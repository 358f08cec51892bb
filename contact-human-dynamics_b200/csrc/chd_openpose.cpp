// OpenPose keypoint ingestion (host side, product code): many `*_keypoints.json` files -> one (n_files, J, 3) fp64 array.
// Replaces the per-file `json.load` + `np.array(...).reshape(-1, 3)` of the reference
// (src/utils/openpose_utils.py:48-66 load_keypoint_file, :68-76 load_keypoint_dir), which is serial Python and dominates the
// end-to-end time of the 100k-window configuration once the classifier itself runs on the GPU.
//
// Semantics kept: only the FIRST person's "pose_keypoints_2d" is used; a frame without people yields J rows of zeros;
// numbers are converted with strtod (correctly rounded, like Python's float()) so the array is bit-identical to the
// reference's.  Files are dealt to worker threads; every worker reads its file in one piece and scans it once.
#include <atomic>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/chd.h"

namespace {

// position just behind the JSON key `"key"` + ':' at or after `p`, nullptr if absent
const char* find_key(const char* p, const char* end, const char* key) {
  const size_t kl = strlen(key);
  for (const char* q = p; q + kl + 2 <= end; ++q) {
    if (*q != '"') continue;
    if (memcmp(q + 1, key, kl) != 0 || q[1 + kl] != '"') continue;
    const char* r = q + kl + 2;
    while (r < end && (*r == ' ' || *r == '\t' || *r == '\n' || *r == '\r')) ++r;
    if (r < end && *r == ':') return r + 1;
  }
  return nullptr;
}
const char* skip_ws(const char* p, const char* end) {
  while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
  return p;
}

// 0 ok, -2 unreadable file, -3 malformed / unexpected keypoint count
int parse_file(const char* path, int J, double* out, std::string& buf) {
  FILE* f = fopen(path, "rb");
  if (!f) return -2;
  buf.clear();
  char chunk[1 << 14];
  size_t got;
  while ((got = fread(chunk, 1, sizeof chunk, f)) > 0) buf.append(chunk, got);
  fclose(f);
  const char* p = buf.data();
  const char* end = p + buf.size();
  const char* q = find_key(p, end, "people");
  if (!q) return -3;
  q = skip_ws(q, end);
  if (q >= end || *q != '[') return -3;
  q = skip_ws(q + 1, end);
  if (q < end && *q == ']') {   // nobody detected: zeros (openpose_utils.py:61-63)
    for (int i = 0; i < 3 * J; ++i) out[i] = 0.0;
    return 0;
  }
  // the first person object: its "pose_keypoints_2d" array (the key must sit before the object closes)
  const char* obj_end = q;
  {
    int depth = 0;
    for (; obj_end < end; ++obj_end) {
      if (*obj_end == '{') ++depth;
      else if (*obj_end == '}' && --depth == 0) break;
    }
  }
  const char* a = find_key(q, obj_end, "pose_keypoints_2d");
  if (!a) return -3;
  a = skip_ws(a, end);
  if (a >= end || *a != '[') return -3;
  ++a;
  int cnt = 0;
  while (true) {
    a = skip_ws(a, end);
    if (a >= end) return -3;
    if (*a == ']') break;
    if (*a == ',') { ++a; continue; }
    char* e = nullptr;
    const double v = strtod(a, &e);   // the buffer is NUL terminated (std::string)
    if (e == a) return -3;
    if (cnt < 3 * J) out[cnt] = v;
    ++cnt;
    a = e;
  }
  return cnt == 3 * J ? 0 : -3;
}

}  // namespace

extern "C" int chd_openpose_load(const char* const* paths, int32_t n_files, int32_t num_joints, double* out, int32_t n_threads) {
  if (!paths || !out || n_files < 0 || num_joints <= 0) return -1;
  if (n_files == 0) return 0;
  int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > n_files) nt = n_files;
  std::atomic<int> next(0), err(0);
  auto work = [&]() {
    std::string buf;
    buf.reserve(1 << 14);
    for (;;) {
      const int i = next.fetch_add(16);
      if (i >= n_files) break;
      for (int k = i; k < i + 16 && k < n_files; ++k) {
        const int rc = parse_file(paths[k], num_joints, out + (size_t)k * num_joints * 3, buf);
        if (rc) {
          int zero = 0;
          err.compare_exchange_strong(zero, rc);
        }
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nt; ++t) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  return err.load();
}
