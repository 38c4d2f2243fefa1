// Host-side plan handle (opaque to C callers) + tile geometry shared by the fused kernels.
#pragma once
#include "cmax_common.cuh"
#include "cmax_tile.cuh"

namespace cmax {

// Source tiles: the plan's sort key is tile-major (32x32 tiles of UN-warped pixels, row-major inside a tile), so that events
// which are neighbours in the stream read neighbouring flow vectors and vote into neighbouring IWE cells.
constexpr int kTile = 32;
constexpr int kRunE = 8;                  // consecutive events one thread of a run kernel walks
constexpr int kWarpTile = 32 * kRunE;    // events per warp-tile of the packed copy
constexpr int kStripTileBytes = 128 + 32 * kRunE * 4;  // 32 strip headers + 32 strips of kRunE times (cmax_events.cu)
constexpr int kStripBinBytes = 32 * kRunE;              // time-aware plans: + one byte per event and reference time (its time bin)

}  // namespace cmax

struct cmax_plan {
  const float* events;  // float4 per event; tile-sorted copy (in the workspace) or the caller's array
  // Private re-packed copy for the run kernels: (x, y, tz, bits(src)) with src = un-warped flat pixel (src/warp.py:305)
  // and tz = normalised dt of reference time 0 when n_ref == 1 (packed_has_dt), else the raw timestamp t.
  // Stored in warp-tile order: event tile*kWarpTile + lane*kRunE + k lives at slot tile*kWarpTile + k*32 + lane, i.e.
  // pre-transposed so that coalesced loads hand every lane kRunE CONSECUTIVE events without a shared-memory pass.
  void* packed;
  int packed_has_dt;
  // compact: every event has integer pixel coordinates < 65536, the packed copy is 8 bytes per event
  // (tz, row<<16|col) in the same warp-tile order; else 16 bytes (x, y, tz, bits(src)).
  int compact_ok;  // eligibility (from validation)
  int compact;     // format currently packed
  // strips: source-pixel runs cut into padded strips of kRunE events (pack_strips_kernel in cmax_events.cu); NULL when
  // the batch does not qualify (not pixel-ordered, fractional coordinates, or too sparse for the padding to pay)
  void* strips;
  int64_t n_strips;
  int strip_tile_bytes;  // kStripTileBytes, + n_ref * kStripBinBytes when the plan was packed for a flow voxel (n_bins > 0)
  const uint32_t* sorted_keys;
  const uint32_t* key_first;   // [n_keys + 1]: first event of every sort key (events of key k = key_first[k] .. key_first[k+1])
  const uint32_t* key_strip0;  // [n_keys + 1]: first strip of every sort key
  int packed_valid;            // the packed copy matches the current reference times / format (built on demand: the strip kernels never read it)
  int64_t n;
  int H, W, pad_h, pad_w, Hp, Wp;
  float t_min, t_max;
  int src_row_lo, src_row_hi;  // rows of the un-warped image that hold events of this batch (a sharded batch that is spatially
                               // compact publishes them, and its peers skip the rows it cannot have touched)
  int order;         // cmax_order
  int vote_variant;  // see cmax_plan_set_variant
  int grad_variant;
  int stage_mask;       // see cmax_plan_set_stage_mask (7 = everything; other values only in CMAX_MEASURE builds)
  int tiles_x, tiles_y, n_tiles;
  cmax_time_params_t* d_params;  // device
  float* d_minmax;               // device [2]
  int32_t* d_status;             // device [1]
  int n_ref, n_bins;
  cmax::TileGeom tile;  // CMAX_MOTION_TILE geometry (cmax_plan_set_tile_flow); tile.hp == 0: not set
  float t_scale;
};
