// One handle, several GPUs, one process (mpcb_settings.n_devices / device_ids).
//
// Every (x0, reference) pair is an independent problem that shares only read-only per-system constants, so a batch is cut
// into contiguous shards -- sizes differ by at most one problem -- that are solved concurrently with no exchange during
// the solve.  The host entries run one host thread per device (each drives its device's stream pair through the ordinary
// single-device path and writes straight into the caller's arrays); the device entry (mpcb_api.cu) fans the shards out and
// the results back over NVLink as peer copies ordered by events.  Helpers shared by mpcb_api.cu and nmpc_api.cu.
#pragma once
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mpcb200.h"
#include "api_common.hpp"

namespace mpcb {

inline void shard_range(long long batch, int r, int n, long long* lo, long long* hi) {
  const long long base = batch / n, extra = batch % n;
  *lo = r * base + std::min<long long>(r, extra);
  *hi = *lo + base + (r < extra ? 1 : 0);
}

struct ShardDims { size_t nx, nu, H, nz, ny; };   // ny: duals per problem (nt for the linear path)

// the slice [lo, lo + n) of a batch io (host or device pointers alike: the layouts are problem-major)
inline mpcb_batch_io shard_io(const mpcb_batch_io& io, long long lo, long long n, const ShardDims& d) {
  mpcb_batch_io s = io;
  const size_t L = (size_t)lo;
  s.batch = n;
  auto off = [&](auto* p, size_t per) { return p ? p + L * per : p; };
  s.x0 = off(io.x0, d.nx);
  s.xref = io.xref_broadcast ? io.xref : off(io.xref, d.nx);
  s.uref = io.uref_broadcast ? io.uref : off(io.uref, d.nu);
  s.warm_u = off(io.warm_u, d.nz); s.warm_y = off(io.warm_y, d.ny);
  s.u = off(io.u, d.nz); s.e_u = off(io.e_u, d.nz);
  s.x = off(io.x, d.nx * (d.H + 1)); s.e_x = off(io.e_x, d.nx * (d.H + 1));
  s.u0 = off(io.u0, d.nu);
  s.status = off(io.status, 1); s.iters = off(io.iters, 1); s.inner_iters = off(io.inner_iters, 1);
  s.prim_res = off(io.prim_res, 1); s.dual_res = off(io.dual_res, 1); s.objective = off(io.objective, 1);
  s.y = off(io.y, d.ny);
  return s;
}

inline mpcb_closed_loop_io shard_closed_loop_io(const mpcb_closed_loop_io& io, long long lo, long long n, size_t nx, size_t nu) {
  mpcb_closed_loop_io s = io;
  const size_t L = (size_t)lo, T = (size_t)io.steps;
  s.batch = n;
  s.x0 = io.x0 + L * nx;
  s.xref = io.xref_broadcast ? io.xref : io.xref + L * nx;
  s.uref = io.uref_broadcast ? io.uref : io.uref + L * nu;
  if (io.x_traj) s.x_traj = io.x_traj + L * nx * (T + 1);
  if (io.u_traj) s.u_traj = io.u_traj + L * nu * T;
  if (io.iters_total) s.iters_total = io.iters_total + L;
  if (io.unsolved_steps) s.unsolved_steps = io.unsolved_steps + L;
  return s;
}

// fn(r) for r = 0 .. n-1, r = 0 on the calling thread and the others on their own host threads (a CUDA context is per device,
// the current device per thread); the first failing shard's code and message are reported through mpcb_last_error().
template <class F>
int run_sharded(int n, F&& fn) {
  std::vector<int> rc((size_t)n, 0);
  std::vector<std::string> msg((size_t)n);
  std::vector<std::thread> th;
  th.reserve((size_t)n);
  for (int r = 1; r < n; r++)
    th.emplace_back([&, r]() { rc[(size_t)r] = fn(r); if (rc[(size_t)r] != MPCB_OK) msg[(size_t)r] = mpcb_last_error(); });
  rc[0] = fn(0);
  if (rc[0] != MPCB_OK) msg[0] = mpcb_last_error();
  for (auto& t : th) t.join();
  for (int r = 0; r < n; r++)
    if (rc[(size_t)r] != MPCB_OK) return api_fail(rc[(size_t)r], "device shard " + std::to_string(r) + ": " + msg[(size_t)r]);
  return MPCB_OK;
}

// n_devices / device_ids of the settings -> list of distinct, valid ordinals (empty: single-device handle)
inline int parse_devices(const mpcb_settings& st, int ndev_visible, std::vector<int>& ids) {
  ids.clear();
  if (st.n_devices <= 1) return MPCB_OK;
  if (st.n_devices > 8) return api_fail(MPCB_ERR_INVALID, "settings.n_devices: at most 8 devices per handle");
  for (int i = 0; i < st.n_devices; i++) {
    const int d = st.device_ids[i];
    if (d < 0 || d >= ndev_visible) return api_fail(MPCB_ERR_INVALID, "settings.device_ids: ordinal out of range");
    if (std::find(ids.begin(), ids.end(), d) != ids.end()) return api_fail(MPCB_ERR_INVALID, "settings.device_ids: duplicate ordinal");
    ids.push_back(d);
  }
  return MPCB_OK;
}

// batches below this many problems per device are not worth a fan-out: the root device takes them alone
constexpr long long MULTI_MIN_PER_DEVICE = 512;

}  // namespace mpcb
