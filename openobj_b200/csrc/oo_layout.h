// Shared constants: parameter-block layout, shared-memory map of one fused tile.
// Plain C++ (no CUDA) so the host-side schedule code and the CPU tile emulator used by the
// tests (tests/emu) see exactly the numbers the kernels use.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define OO_HOSTDEV __host__ __device__
#else
#define OO_HOSTDEV
#endif

namespace oo {

// ---- model dimensions (reference: trainer.py:20-21, model.py:31-56, room_0.json) -------------
constexpr int H = 32;          // hidden width
constexpr int NDIR = 21;       // projection directions (embedding.py:15-37)
constexpr int NBAND = 6;       // 2^0..2^5
constexpr int E1 = 87;         // 3 + 21*4
constexpr int E2 = 42;         // 21*2
constexpr int EMB = 129;
constexpr int C = 512;         // clip / part feature width
constexpr int S = 10;          // samples per ray

// ---- parameter block (floats per object); tensor order = named_parameters() ------------------
constexpr int NT = 19;
constexpr int OFF_IN_W = 0,      SZ_IN_W = H * E1;          // [32][87]
constexpr int OFF_IN_B = 2784;
constexpr int OFF_M1_W = 2816,   SZ_M1_W = H * H;
constexpr int OFF_M1_B = 3840;
constexpr int OFF_CAT_W = 3872,  SZ_CAT_W = H * (H + E1);   // [32][119]
constexpr int OFF_CAT_B = 7680;
constexpr int OFF_M2_W = 7712;
constexpr int OFF_M2_B = 8736;
constexpr int OFF_A_W = 8768;                                // [1][32]
constexpr int OFF_A_B = 8800;                                // [1]
constexpr int OFF_CL_W = 8804,   SZ_CL_W = H * (H + E2);    // [32][74]
constexpr int OFF_CL_B = 11172;
constexpr int OFF_OC_W = 11204;                              // [3][32]
constexpr int OFF_OC_B = 11300;                              // [3]
constexpr int OFF_CP_W = 11304;                              // [32][74]
constexpr int OFF_CP_B = 13672;
constexpr int OFF_OCL_W = 13704, SZ_OCL_W = C * H;          // [512][32]
constexpr int OFF_OCL_B = 30088;                             // [512]
constexpr int OFF_PE_B = 30600;                              // [21][3]
constexpr int PEND = 30664;
constexpr int PSTRIDE = 30720;
constexpr int PCOUNT = 30659;

constexpr int kOff[NT] = {OFF_IN_W, OFF_IN_B, OFF_M1_W, OFF_M1_B, OFF_CAT_W, OFF_CAT_B, OFF_M2_W, OFF_M2_B,
                          OFF_A_W, OFF_A_B, OFF_CL_W, OFF_CL_B, OFF_OC_W, OFF_OC_B, OFF_CP_W, OFF_CP_B,
                          OFF_OCL_W, OFF_OCL_B, OFF_PE_B};
constexpr int kSize[NT] = {SZ_IN_W, H, SZ_M1_W, H, SZ_CAT_W, H, H * H, H, H, 1, SZ_CL_W, H, 3 * H, 3,
                           SZ_CL_W, H, SZ_OCL_W, C, NDIR * 3};
// AdamW parameter groups (which tensors autograd reaches)
constexpr int GROUP_TRUNK = 0, GROUP_COLOR = 1, GROUP_CLIP = 2;
OO_HOSTDEV inline constexpr int group_of_offset(int off) {
    return off < OFF_CL_W ? GROUP_TRUNK : off < OFF_CP_W ? GROUP_COLOR : off < OFF_PE_B ? GROUP_CLIP : GROUP_TRUNK;
}

// ---- one tile: RT rays x S samples = P points, activations feature-major [row][P] -----------
constexpr int RT = 10;
constexpr int P = RT * S;      // 100
constexpr int PS = 100;        // row stride (floats); PS % 32 == 4: the tensor-core fragment loads (8 rows x 4 points, or
                               // 4 row pairs x 8 points) touch 32 distinct banks
constexpr int NTHREADS = 512;      // 16 warps: four per scheduler hide the tensor-pipe and shared-memory latencies
constexpr int NWARPS = NTHREADS / 32;

// activation rows
constexpr int R_H2 = 0;        // fc2            } contiguous = cat_layer input [h2 ; e1]
constexpr int R_E1 = 32;       // e1 (87) + 1 zero row
constexpr int E1P = 88;
constexpr int R_H4 = 120;      // fc4            } contiguous = head input [h4 ; e2]
constexpr int R_E2 = 152;      // e2 (42) + 6 zero rows (head input padded to 80 = ten 8-wide k-steps)
constexpr int E2P = 48;
constexpr int R_H1 = 200;
constexpr int R_H3 = 232;
constexpr int R_HC = 264;      // color_linear out } contiguous = [hc ; hp]
constexpr int R_HP = 296;      // clip_linear out
constexpr int R_T = 328;       // scaled coords t (3) + 1 spare
constexpr int R_MISC = 332;    // 12 rows, see M_*
constexpr int NROWS = 344;
constexpr int M_OCC = 0, M_TERM = 1, M_DRAW = 2, M_COL = 3, M_DCOL = 6, M_FREE = 9, M_HU = 10, M_Z = 11;   // M_Z: z of the tile's points

constexpr int SM_ACT = 0;
constexpr int SM_W = NROWS * PS;                     // 34400
// padded weight copies (zero pad columns); K padded to multiples of 8 = one m16n8k8 k-step
constexpr int WS_IN = 88, WS_H = 36, WS_CAT = 120, WS_HD = 80;
constexpr int KP_IN = 88, KP_CAT = 120, KP_HD = 80;
constexpr int W_IN = 0;
constexpr int W_M1 = W_IN + H * WS_IN;               // 2816
constexpr int W_CAT = W_M1 + H * WS_H;               // 3968
constexpr int W_M2 = W_CAT + H * WS_CAT;             // 7808
constexpr int W_CL = W_M2 + H * WS_H;                // 8960   } contiguous [64][80]
constexpr int W_CP = W_CL + H * WS_HD;               // 11520
constexpr int W_A = W_CP + H * WS_HD;                // 14080
constexpr int W_OC = W_A + H;                        // 14112
constexpr int B_IN = W_OC + 3 * H;                   // 14208
constexpr int B_M1 = B_IN + H, B_CAT = B_M1 + H, B_M2 = B_CAT + H, B_CL = B_M2 + H, B_CP = B_CL + H;
constexpr int B_A = B_CP + H;                        // 14400
constexpr int B_OC = B_A + 4;
constexpr int W_PE = B_OC + 4;                       // 14408, [21][3]
constexpr int W_TOTAL = W_PE + 64;                   // 14472
constexpr int SM_RAY = SM_W + W_TOTAL;               // 48872
constexpr int RP = 12;                               // padded rays per row
constexpr int SM_ST = SM_RAY;                        // S^T [32][12]   sum_i T_i hp_i
constexpr int SM_UT = SM_ST + H * RP;                // U^T [32][12]   dL/dS
constexpr int SM_RV = SM_UT + H * RP;                // ray values [24][12]
constexpr int NRV = 27;                              // rows 0..23 per-ray values (zeroed with the tile), 24..26 per-object wb / bb
constexpr int NRV_TILE = 24;
constexpr int SM_UPART = SM_RV + NRV * RP;           // gs [32][12] = G S + opac wb (phase 12)
constexpr int SM_FEAT = SM_UPART + H * RP;           // Y [10][512]: gt part features of the tile's rays (cp.async in phase 0);
                                                     // phase 10 overwrites each warp's 32 columns with its partial W^T y
constexpr int YSTR = C + 4;                          // row stride of Y: YSTR % 32 == 4 -> conflict-free A fragments in phase 10
constexpr int SM_YS = SM_FEAT + RT * YSTR;              // gt-feature statistics: partials [10][16 warps][2], totals [2][12]
constexpr int SM_FROW = SM_YS + 352;                 // int [12]: feature-table rows of the tile's rays (phase 0 -> phase 1)
constexpr int SM_G = SM_FROW + 12;                   // G = W_ocl^T W_ocl [32][32] of the current object (stage_derived)
constexpr int SM_TOTAL = SM_G + H * H;               // floats
// ray-value rows
constexpr int V_DEPTH = 0, V_OPAC = 1, V_COL = 2, V_GD = 5, V_GO = 6, V_GC = 7, V_CF = 10, V_BG = 11,
              V_LD = 12, V_LC = 13, V_LO = 14, V_LF = 15, V_A = 16, V_B = 17, V_ZSRC = 18,
              V_GTD = 19, V_LAB = 20, V_RGB = 21,      // gt depth, label, gt colour of the tile's rays (phase 0)
              V_WB = 24;                               // W_ocl^T b_ocl [32] and b.b [1] of the current object (gram_stage)

static_assert(SM_TOTAL * 4 <= 232448, "tile does not fit in 227 KB of shared memory");
static_assert(PS % 4 == 0 && SM_W % 4 == 0 && SM_RAY % 4 == 0 && SM_FEAT % 4 == 0 && SM_G % 4 == 0, "float4 alignment");
// owner threads of the small accumulators, spread so that no warp carries more than one of the side jobs of the fused
// phase 42: M (256 threads from M_T0), out_color / out_alpha weights (128 threads from OC_T0), m / beta (threads 0..32)
constexpr int M_T0 = 128, OC_T0 = 384;
static_assert(OC_T0 + 4 * H <= NTHREADS && M_T0 + 8 * H <= OC_T0, "owner-thread ranges");

// per-object constants derived from the out_clip layer (k_gram): G = W^T W [32][32], wb = W^T b [32], bb = b.b
constexpr int DER_G = 0, DER_WB = 1024, DER_BB = 1056, DERIVED = 1088;
// per-ray record K1 leaves for K4 (out_clip gradient): A, opacity, S[32]
constexpr int REC_A = 0, REC_OPAC = 1, REC_S = 4, RAYREC = 36;   // S 16-byte aligned
// out_clip region of a gradient slab holds M = sum_r B_r S_r S_r^T [32][32], m = sum_r B_r opac_r S_r [32], beta = sum_r B_r opac_r^2
constexpr int SLAB_M = OFF_OCL_W, SLAB_MV = OFF_OCL_W + 1024, SLAB_BETA = OFF_OCL_W + 1056;

constexpr float PI_F = 3.14159274101257324f;         // float32(np.pi): embedding.py:52 multiplies in fp32

// p-ranges of the four point groups used by the 32x32 weight-gradient tiles
OO_HOSTDEV inline constexpr int pgroup_begin(int g) { return g == 0 ? 0 : 28 + 24 * (g - 1); }
OO_HOSTDEV inline constexpr int pgroup_end(int g) { return 28 + 24 * g; }

// ---- static schedule layout (ints) -----------------------------------------------------------
// sched[0 .. n_cta]            : first global tile (object * tiles_per_obj + tile) of each CTA, sentinel at n_cta
// sched[n_cta+1 .. 2 n_cta]    : first slot id of each CTA
// sched[2 n_cta+1 .. +n_obj+1] : slot range [begin, end) per object (prefix array, n_obj+1 entries)
inline int sched_ints(int n_cta, int n_obj) { return 2 * n_cta + 1 + n_obj + 1; }

}  // namespace oo
