// Tile-layout recurrent kernel of the H = 384 LSTM (see lstm_rec_tc6_impl.cuh for the design), instantiated for two tile
// shapes; B200_LSTM_SHAPE selects one for the whole process (the engine asks b200_lstm_tile_chunks once per plan):
//   3x16  three interleaved sub-tiles of 16 chunks = 48-chunk tiles (11 clusters of 6 CTAs for a 512-chunk batch)
//   2x32  two interleaved sub-tiles of 32 chunks   = 64-chunk tiles ( 8 clusters): half as many tcgen05.mma per chunk (the
//         issuing warp's ~900 cycles per sub-tile and step bound the 3x16 step time) and half as many exchange operations
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

#define LSTM6_NAMESPACE lstm6_3x16
#define LSTM6_NS 3
#define LSTM6_SN 16
#define LSTM6_GXD 4
#include "lstm_rec_tc6_impl.cuh"
#undef LSTM6_NAMESPACE
#undef LSTM6_NS
#undef LSTM6_SN
#undef LSTM6_GXD

#define LSTM6_NAMESPACE lstm6_2x32
#define LSTM6_NS 2
#define LSTM6_SN 32
#define LSTM6_GXD 3
#include "lstm_rec_tc6_impl.cuh"
#undef LSTM6_NAMESPACE
#undef LSTM6_NS
#undef LSTM6_SN
#undef LSTM6_GXD

static int shape_2x32() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("B200_LSTM_SHAPE");
        mode = (e && strcmp(e, "2x32") == 0) ? 1 : 0;
    }
    return mode;
}

int lstm_rec_tile_chunks(int hidden) { return hidden != 384 ? 0 : (shape_2x32() ? lstm6_2x32::NB : lstm6_3x16::NB); }
int lstm_rec_tile_cluster(int hidden) { return hidden == 384 ? 6 : 0; }
size_t lstm_rec_tile_workspace_bytes(int N) { return shape_2x32() ? lstm6_2x32::workspace_bytes(N) : lstm6_3x16::workspace_bytes(N); }

// gx [tiles][T][6][NB][256], y [tiles][T][NB][H]; tiles = ceil(N / NB), the last one may be partial;
// workspace: lstm_rec_tile_workspace_bytes(N) bytes of exchange staging (contents irrelevant)
int launch_lstm_rec_tc6(const __half* gx, const __half* whh, __half* y, void* workspace, int T, int N, int hidden,
                        int reverse, cudaStream_t stream) {
    return shape_2x32() ? lstm6_2x32::launch(gx, whh, y, workspace, T, N, hidden, reverse, stream)
                        : lstm6_3x16::launch(gx, whh, y, workspace, T, N, hidden, reverse, stream);
}

int copy_lstm_timeline6(long long* host_out, int max_steps) {
    return shape_2x32() ? lstm6_2x32::copy_timeline(host_out, max_steps) : lstm6_3x16::copy_timeline(host_out, max_steps);
}
