/*
 * jmc_kernels.cuh -- the sm_100a kernels of the decoded-surface format path (umbrella header).
 *
 * Everything here is HBM-bound byte movement (arithmetic intensity ~0; ~6 int-ops/B for the colour
 * kernels), so the design rules are: the copy engine (cp.async.bulk) or 16-byte coalesced vector accesses,
 * every byte of a tile in flight before the first store, fixed-size tiles with one CTA per tile (measured
 * faster than a persistent loop), ONE launch per batch of frames, no tensor cores.
 *
 *   jmc_k_common.cuh  shared types and helpers
 *   jmc_k_planes.cuh  planes_kernel, bulk_planes_kernel       NV12 <-> tight NV12 / I420, addressable rows
 *                     (replace nv_dec/nv_dec.cpp:782-820, intel_dec/intel_dec.cpp:284-314,
 *                      intel_enc/intel_enc.cpp:291-307,366-380, the InterleaveUV launch of nv_enc/nv_enc.cpp:1041-1081)
 *   jmc_k_rows.cuh    rows_kernel, bulk_rows_kernel, bulk_rows_pack_kernel   the same ops for widths % 16 != 0
 *   jmc_k_rgb.cuh     rgb_kernel, rgb_bulk_kernel, rgb_to_nv12_kernel        integer BT.601 colour conversion
 *
 * Which kernel a job gets is decided on the host, per launch, in jmc_kernels.cu.
 */
#pragma once
#include "jmc_k_common.cuh"
#include "jmc_k_planes.cuh"
#include "jmc_k_rows.cuh"
#include "jmc_k_rgb.cuh"
