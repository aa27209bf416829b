// pose.cu -- the pose-guided adjacency on the GPU.
//
// Reference: torchreid/dataset_loader.py, generate_graph :218-343 and adj_graph :345-388, canonical configuration
// (num_parts = 3, num_split = 4 with pyramid strips [4,2,1] -> P = 7, method 'same', num_scale = 1).  Per frame and
// body-part class (head / body / leg, keypoint ids :318-320) the reference collects the strips that hold a keypoint
// with confidence > threshold (:323-328, bisect_right over np.arange(0, h + 1, h / 4)), makes the set contiguous
// (:329-333), adds the coarser pyramid strips (:356-371), and links every ordered pair of distinct nodes that share
// a class anywhere in the tracklet (:373-387, itertools.permutations in a loader worker).
//
// The graph is therefore fully described by THREE V-bit membership masks per tracklet (V = S * 7 <= 64): 24 bytes
// instead of the 12.5 KB fp32 matrix.  This file builds the masks from the raw detections (one warp per tracklet),
// expands them to the reference's dense matrix when a caller wants it, and head.cu consumes the masks directly
// (agrl_head_forward_compact_dev) so the dense matrix never has to exist.
#include "common.cuh"

namespace agrl {

__constant__ int kPartOf[18] = {0, 0, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 0, 0, 0, 0};   // dataset_loader.py:318-320

__global__ void pose_masks_kernel(const double *__restrict__ kp, const double *__restrict__ heights,
                                  const uint8_t *__restrict__ valid, int64_t batch, int S, double threshold,
                                  uint64_t *__restrict__ masks) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
    if (b >= batch) return;                                   // whole warps only (blockDim is a multiple of 32)
    uint32_t sets[3] = {0u, 0u, 0u};                          // bit q: quarter strip q + 1 holds the part
    if (lane < S && (valid == nullptr || valid[b * S + lane])) {
        const double h = heights[b * S + lane];
        const double step = h / 4.0;                          // np.arange(0, h + 1, h / num_split)   (:313)
        if (step > 0.0) {                                     // h == 0 raises in numpy -> bare except -> empty sets
            const int n = static_cast<int>(ceil((h + 1.0) / step));      // numpy's arange length
            const double *f = kp + (b * S + lane) * 54;
            for (int p = 0; p < 18; ++p) {
                if (f[3 * p + 2] > threshold) {               // :323
                    const double y = f[3 * p + 1];
                    int loc = 0;                               // bisect_right: entries NOT greater than y (NaN -> all)
                    for (int i = 0; i < n; ++i) loc += !(y < static_cast<double>(i) * step);
                    loc = min(4, max(1, loc));                 // :327
                    sets[kPartOf[p]] |= 1u << (loc - 1);
                }
            }
        }
    }
    uint64_t out[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        uint32_t q = sets[c];
        if (q) {
            const int lo = __ffs(q) - 1, hi = 31 - __clz(q);
            q = ((2u << hi) - 1u) & ~((1u << lo) - 1u);       // contiguous range (:329-333)
            q |= ((q & 3u) ? 16u : 0u) | ((q & 12u) ? 32u : 0u) | 64u;   // halves 5, 6 and the whole strip 7 (:364-365)
        }
        const uint64_t m = (lane < S) ? static_cast<uint64_t>(q) << (7 * lane) : 0ull;
        const uint32_t lo32 = __reduce_or_sync(0xffffffffu, static_cast<uint32_t>(m));
        const uint32_t hi32 = __reduce_or_sync(0xffffffffu, static_cast<uint32_t>(m >> 32));
        out[c] = (static_cast<uint64_t>(hi32) << 32) | lo32;
    }
    if (lane < 3) masks[b * 3 + lane] = (lane == 0) ? out[0] : (lane == 1 ? out[1] : out[2]);
}

__global__ void adj_expand_kernel(const uint64_t *__restrict__ masks, int V, float *__restrict__ adj) {
    const int64_t b = blockIdx.x;
    const uint64_t m0 = masks[b * 3], m1 = masks[b * 3 + 1], m2 = masks[b * 3 + 2];
    float *dst = adj + b * V * V;
    for (int i = threadIdx.x; i < V * V; i += blockDim.x) {
        const int r = i / V, c = i % V;
        const uint64_t hit = ((m0 >> r) & (m0 >> c)) | ((m1 >> r) & (m1 >> c)) | ((m2 >> r) & (m2 >> c));
        dst[i] = (r != c && (hit & 1ull)) ? 1.0f : 0.0f;      // permutations(): distinct pairs only (:386-387)
    }
}

}  // namespace agrl

using namespace agrl;

extern "C" int agrl_pose_part_masks_dev(const double *keypoints, const double *heights, const uint8_t *valid,
                                        int64_t batch, int32_t seq_len, int32_t num_split, double threshold,
                                        uint64_t *masks, void *stream) {
    if (!keypoints || !heights || !masks || batch < 0 || seq_len < 1) return AGRL_E_INVALID;
    if (num_split != 4 || seq_len * 7 > 64) return AGRL_E_UNSUPPORTED;
    int rc = agrl_device_ok();
    if (rc) return rc;
    if (batch == 0) return AGRL_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int warps_per_block = 8;
    const int64_t blocks = (batch + warps_per_block - 1) / warps_per_block;
    pose_masks_kernel<<<static_cast<unsigned>(blocks), 32 * warps_per_block, 0, st>>>(keypoints, heights, valid, batch, seq_len,
                                                                                      threshold, masks);
    AGRL_LAUNCH_CHECK(st, "pose_masks");
    return AGRL_OK;
}

extern "C" int agrl_pose_adjacency_dev(const uint64_t *masks, int64_t batch, int32_t nodes, float *adj, void *stream) {
    if (!masks || !adj || batch < 0 || nodes < 1) return AGRL_E_INVALID;
    if (nodes > 64) return AGRL_E_UNSUPPORTED;
    int rc = agrl_device_ok();
    if (rc) return rc;
    if (batch == 0) return AGRL_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    adj_expand_kernel<<<static_cast<unsigned>(batch), 256, 0, st>>>(masks, nodes, adj);
    AGRL_LAUNCH_CHECK(st, "adj_expand");
    return AGRL_OK;
}
