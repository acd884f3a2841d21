// dev_types.h -- device-side data layout of one BA window (see DESIGN.md "Data layout in HBM").
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace cmlba {

constexpr int MAXF = 16;          // CMLBA_MAX_FRAMES
constexpr int RJ_STRIDE = 36;     // floats per residual Jacobian record (x[10] y[10] A[3] B[6] BR[6] pad)
constexpr int T_STRIDE = 16;      // floats per (point,target) Schur row: JpJdF[8] bd Hdd Hcd[4] good pad
constexpr int ACC_N = 96;         // 91 unique entries of the 13x13 block, padded to 3*32
constexpr int ACC_CHUNK = 32;     // residuals per linearize warp pass (one thread per residual; consecutive residuals of the tile-sorted order)
// linearize_tile_kernel: target-image tiles staged by TMA (cp.async.bulk.tensor.2d) into a shared-memory ring
constexpr int LT_TILE_W = 64, LT_TILE_H = 32;         // core tile of the target image that owns a residual (by its centre projection at bin time)
constexpr int LT_HALO = 3;                            // pattern reach (2) + bilinear tap (1); footprints that drift out of the box since the binning read global memory
constexpr int LT_BOX_W = LT_TILE_W + 2 * LT_HALO;     // 70 texels
constexpr int LT_BOX_H = LT_TILE_H + 2 * LT_HALO;     // 38 rows
constexpr int LT_TILE_BYTES = LT_BOX_W * LT_BOX_H * 16;   // 42 560 B of float4 texels per staged tile
constexpr int LT_STAGE_STRIDE = (LT_TILE_BYTES + 127) & ~127;   // TMA destinations are 128-byte aligned
// ring depth and consumer warps per CTA (+1 producer warp) are template parameters of the kernel (Engine::launch_linearize picks the variant)
#ifndef CMLBA_ACC_SLICES
#define CMLBA_ACC_SLICES 16
#endif
constexpr int ACC_SLICES = CMLBA_ACC_SLICES;                        // CTAs per (host,target) bin that evaluate addToHessianTop from the Jacobian records
constexpr int P2P_POST_CAND_MAX = 65536;      // candidates per rank record that fit the peer-memory exchange (else NCCL all-gather)
constexpr size_t P2P_POST_DOUBLES = 8 + P2P_POST_CAND_MAX / 2;
constexpr size_t P2P_SLOT_DOUBLES = 2 * (size_t) (8 * MAXF + 4) * (8 * MAXF + 4) + 2 * (8 * MAXF + 4);   // sys at the largest window
#ifndef CMLBA_SC_CHUNK
#define CMLBA_SC_CHUNK 64
#endif
constexpr int SC_CHUNK = CMLBA_SC_CHUNK;      // points per Schur CTA (all hosted in one frame)

enum : uint8_t { RES_IN = 0, RES_OOB = 1, RES_OUTLIER = 2 };

// per (host h, target t) constants, index h*N+t.  DSOFramePrecomputed (DSOFrame.h:248-291)
struct PairPre {
    double R[9], t[3];    // current estimate  PRE_worldToCam_t * PRE_camToWorld_h  (trialRefToTarget)
    double R0[9], t0[3];  // at the FEJ evaluation points (PRE_RTll_0 / PRE_tTll_0)
    double a, b;          // exposure transition aff_h.to(aff_t)  (map/Exposure.h:119-123)
    float b0;             // hostData->getB0(scaleB)  (DSOFrame.h:197-199)
    float pad;
    float Af[3], Bf[3];   // R[:,0]/fx, R[:,1]/fy: P(x+sx, y+sy) = P(x,y) + sx*A + sy*B  (pattern offsets in fp32, see linearize.cuh)
};

// DSOFrame (DSOFrame.h:17-246) as a POD
struct FrameDev {
    double evalR[9], evalt[3];   // worldToCam_evalPT
    double preR[9], pret[3];     // PRE_worldToCam
    double state[10], state_zero[10], state_backup[10], state_scaled[10], step[10];
    double prior[8];
    double exposure;             // ab_exposure
    float energy_th;             // frameEnergyTH
    int keyid;
};

struct Ctrl {
    int cur;            // index of the committed buffer set (0/1); the other one holds the candidate linearization
    int done;           // GN loop finished (canbreak && it>=1): later kernels of a pre-recorded run become no-ops
    int canbreak;
    int failed;         // non-finite energy / step
    int iteration;      // GN iterations completed
    int accepted;
    int num_dropped;
    int pad0;           // set when a peer-memory exchange timed out (reported as CMLBA_ERR_STATE, not as a numeric failure)
    double lambda;
    double energy_last;     // energy of the committed linearization
    double energy_new;      // energy of the candidate linearization
    double energy_first;
    float sumA, sumB, sumT, sumR;   // doStepFromBackup accumulators (BA:950-1018)
    double sumNID;
    int numID;
    int sc_done_count;      // last-block counter (point step kernel)
    double stats[16];       // ||HA|| ||Hsc|| ||bA|| ||bsc|| ||x|| ... (the Statistic series of BA.h:215-233)
    // step rejection (forceAccept = false, BA:843-877): linearized energies of calcLEnergy (BA:2118-2208)
    double energyL_last, energyL_new;
    double prior_energy_pts;    // sum_p deltaF^2 * priorF (BA:2200) of the current point states
    int rejected_at;            // value of `iteration` right after the last rejected step (restore_state_kernel keys on it)
    int rejected;               // number of rejected steps
    int pt_bad;                 // non-finite point steps of this rank (point_step_kernel)
    int asm_done_count;         // last-block counter of assemble_kernel (peer signalling)
    int acc_done_count;         // jobs of accumulate_kernel finished since prepare(): stitch_pair_kernel waits on it (the kernels run on two streams, no stream-level join)
    int pad1;                   // set when stitch_pair_kernel's bounded wait for acc_done_count expired (reported as CMLBA_ERR_STATE)
    int final_done;             // the closing linearizeAll(true) of run() has been executed (guard 2 of the pre-launched closing sequence)
    int pad2;
};

struct DevWin {
    // sizes
    int N, P, R, W, H, n;          // n = 8N+4
    int newest_begin;              // residuals [newest_begin, R) target the newest frame (sorted by bin t*N+h)
    int n_sc_chunks;
    // calibration / parameters
    double fx, fy, cx, cy, fxi, fyi;
    float huber, cth, scaleF, scaleC, scaleA, scaleB, scaleT, scaleR;
    float th_opt;
    int optA, optB, force_accept, fix_lambda, idepth_fix_prior;
    double fixed_lambda;
    // frames
    const float4 *img[MAXF];       // level-0 (I,dx,dy,0) texels, row-major
    FrameDev *frames;
    PairPre *pairs;
    Ctrl *ctrl;
    const double *AH, *AT;         // [h*N+t][8][8] adjoints (BA:1071-1095)
    const double *HM, *bM;         // marginalisation prior (n x n, n) or zero
    int has_HM;                    // 0: the prior is all zero (disableMarginalization, BA:1395-1398) and is skipped
    const double *Pns;             // nullspace projector 0.5(NN+^T + ...) (BA:1247-1249), n x n
    // points (sorted by host frame)
    const int *pt_host;
    const float *pt_x, *pt_y;
    double *pt_idepth;
    float *pt_idepth_zero, *pt_idepth_backup;
    const float *pt_colors, *pt_weights;   // [P][8]
    const float *pt_priorF;
    float *pt_Hdd, *pt_bd, *pt_Hcd, *pt_HdiF, *pt_bdSumF, *pt_idepth_hessian, *pt_max_rel_bs;
    int *pt_num_good;              // numGoodResiduals
    int *pt_ngood_cur;             // good residuals in the committed linearization
    double *pt_step;
    // residuals as the host lays them out (bin-major: bin = t*N+h, then by device point): inputs of the tile binning only
    const int *r_point;
    const uint8_t *r_host, *r_target;
    const int *res_bin_begin;      // [N*N+1] first host-order residual of every bin
    // tile binning, rebuilt on the device by every prepare() (bin_count / bin_scatter / bin_segments kernels, linearize.cuh):
    // residuals sorted by (target, tile of the centre projection, host); every per-residual array below is in THIS order
    int tiles_x, tiles_y, n_tiles; // LT_TILE_W x LT_TILE_H tiles per frame
    int n_chunks;                  // ceil(R / 32) warp passes of linearize_tile_kernel
    int tma_on;                    // tensor maps encoded: tiles are staged by TMA (else every tap is read from global memory)
    int lt_grid;                   // CTAs of linearize_tile_kernel (cta_info is laid out for this grid)
    float4 *r_pt4;                 // [5][R] per-residual copies of the point constants in the sorted order: (x, y, -, -), colours[0..3], [4..7], weights[0..3], [4..7]
    int *cta_info;                 // [lt_grid][LT_INFO_INTS] q0, q1, first target, -, descriptors[4], users[4] of the first tiles, tile table (linearize.cuh)
    long long *lt_trace;           // development (lt_mode & 2): [lt_grid][16 warps][32] SM clock stamps of the first passes
    unsigned long long *ktrace_base;   // first slot of the timeline buffer (phase stamps of solve_kernel live behind the launch slots)
    unsigned long long *ktrace;    // development (CMLBA_KTRACE): [kernel][3] globaltimer of first CTA scheduled / first CTA past its dependency wait / last warp done
    int lt_mode;                   // development: 1 = the consumers only run the ring protocol (TMA streaming floor of the pass)
    int lt_exact;                  // 1 = every pattern pixel is projected in fp64 like the reference (parity study; default: fp32 offsets from the fp64 centre)
    int *bin_key;                  // [R] host order: ((t * n_tiles + tile) * N + h)
    uint8_t *bin_g;                // [R] host order: shared-memory bank group of the centre texel inside its tile box (bin_interleave_kernel)
    int *bin_hist;                 // [N * n_tiles * N] zero between uses
    int *bin_offs;                 // [N * n_tiles * N] exclusive scan of the histogram
    int *job_of_tile;              // [N * n_tiles] index of the (t, tile) job among the non-empty ones
    uint32_t *job_desc;            // [jobs] t | tile_x << 4 | tile_y << 16
    int *job_begin;                // [jobs + 1] first sorted residual of every tile job
    uint32_t *r_pht;               // [R] device point | host << 24 | target << 28
    int *r_job;                    // [R] tile job of the residual
    int *r_src;                    // [R] host-order index of the residual
    int *bin_ticket;               // last-block counters of the binning kernels
    // final states in HOST order, written by the fixLinearization pass (what finish_run reads back)
    uint8_t *fin_state, *fin_alive;
    float *fin_energy;
    uint8_t *r_state[2];
    float *r_energy[2];
    uint8_t *r_good[2];
    uint8_t *r_new_state;
    float *r_new_energy, *r_new_energy_wo;
    uint8_t *r_alive;
    float *r_center;               // [R][3] centerProjectedTo
    float *rj[2];                  // [R][RJ_STRIDE] Jacobian records (committed / candidate) in the HOST's residual order; rec[35] = 1 marks a good residual
    float *T[2];                   // [P][N][T_STRIDE]
    float *dbg;                    // optional [R][40]: resF[8] JIdx[16] JabF[16]
    // partial sums
    double *energy_part;           // [n_chunks]
    int acc_target;                // value acc_done_count reaches when the accumulation this stitch depends on has finished (0: no device-side wait)
    float *acc_bin;                // [N*N (bin = t*N+h)][ACC_SLICES][ACC_N] 13x13 blocks of the committed linearization (accumulate_kernel)
    float *sc_part;                // [n_sc_chunks][sc_stride]
    int sc_stride;                 // (8N)^2 + 32N + 8N + 16 + 4 (padded to 4)
    const int *sc_chunk_host, *sc_chunk_begin, *sc_chunk_count;
    const int *host_chunk_begin;   // [N+1]
    double *st_out;                // [N*N][st_stride(N)] block products of every ordered frame pair (stitch_pair_kernel)
    double *sys;                   // summed system: HA[n*n] bA[n] HS[n*n] bS[n]  (allreduce payload)
    double *x;                     // [n]
    double *xAd;                   // [h*N+t][8]
    double *pt_part;               // point-step partial sums [blocks][2]
    int n_pt_blocks;
    int update_points_only;
    int marg_mode;                 // 1: marginalizePointsF pass (BA:2466-2513): residual vectors res_toZeroF = resF - J*delta (fixLinearization, BA:2210-2238), no prior shift in the Schur step
    const float *pair_delta;       // [h*N+t][8] adHTdeltaF (computeDelta, BA:1105-1117)
    // multi-GPU (points sharded, frames replicated): every linearization all-gathers one record per rank
    //   [energy, sumNID, numID, bad, prior_energy, 0, 0, 0 (doubles) | cand_cap candidate energies of the newest frame (floats, -1 = none)]
    int world, rank, cand_cap;
    // peer exchange of the reduced system over NVLink (cudaIpc-mapped buffers, see p2p_reduce in kernels.cuh); null = NCCL path
    char *p2p_base[MAXF];          // every rank's exchange buffer: [flags u64[16] | pad to 256 B | slot 0 | slot 1]
    unsigned long long p2p_epoch;  // number of this exchange (same on every rank); slot = epoch & 1
    unsigned long long p2p_post_epoch;   // same for the post-linearize records (flags u64[16] at byte 128)
    int p2p_on, p2p_post_on;
    int post_stride;               // doubles between the ranks' records in post_recv
    double *post_send;             // this rank's record
    const double *post_recv;       // world records
};

}  // namespace cmlba
