// Internal declarations shared by the translation units of libcodecad_b200.
#ifndef CC_INTERNAL_H
#define CC_INTERNAL_H

#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/codecad_b200.h"
#include "cc_microcode.h"

struct cc_decoded {
    std::vector<uint32_t> microcode;
    cc_program_info info;
};

// cc_program.cpp
int cc_decode_program(const float *words, uint32_t n_words, cc_decoded *out, std::string *err);

struct cc_program {
    cc_decoded dec;
    uint32_t *d_code;  // device copy of the microcode
    uint64_t id;       // identifies what is currently loaded in the __constant__ window
};

// ---- kernel launch layer (cc_kernels.cu) ------------------------------------------------

enum cc_sink_kind { CC_SINK_FLOAT4 = 0, CC_SINK_PYMCUBES, CC_SINK_CLASSIFY, CC_SINK_MASS };

struct cc_block_desc {  // one block of a subdivision level (device resident)
    float cx, cy, cz;   // fp32 corner of the first sample (cell centre), reference rounding
    uint32_t pad;
};

struct cc_eval_args {
    const uint32_t *code;
    uint32_t code_words;
    uint32_t n_slots;
    // geometry: single grid (blocks == nullptr) or a list of equally sized blocks
    float cx, cy, cz, step;
    uint32_t nx, ny, nz, x_offset;
    uint32_t n_blocks;
    uint32_t tiles_per_block;
    const cc_block_desc *blocks;
    // sinks
    void *out;            // float4* / float*            (FLOAT4, PYMCUBES)
    float threshold;      // CLASSIFY, MASS
    uint32_t *counter;    // running length of `list`   (CLASSIFY, MASS)
    uint8_t *list;        // uchar4 (x,y,z,0) per hit, or with `list_block` the hierarchy form
    uint32_t *list_block; // optional: block index of every hit (hierarchy fast path)
    uint32_t *sums;       // MASS: 10 uint32 per block (stride 10), or one set when blocks == nullptr
    // decoupled look-back scratch (ordered compaction)
    uint32_t *ticket;
    unsigned long long *tile_status;
};

struct cc_launch_cfg {
    int pts;         // points per thread: 1, 2, 4
    int prog_space;  // 1 = __constant__, 2 = shared
};

// launches on `stream`; returns cudaError_t as int
int cc_launch_eval(int sink, const cc_launch_cfg &cfg, const cc_eval_args &args, void *stream);
int cc_upload_constant_program(const uint32_t *h_code, uint32_t n_words, void *stream);
uint32_t cc_tile_points(const cc_launch_cfg &cfg);
size_t cc_eval_smem_bytes(const cc_launch_cfg &cfg, uint32_t n_slots, uint32_t code_words);

// hierarchy helper kernels
struct cc_level_geom {
    double ox, oy, oz;      // origin
    double resolution;
    double half_cell;       // int_step / 2 (leaf units) added to the int corner; 0 on flat axes
    int dimension;
};
// subdivision: corner = fp32((int_corner + int_step/2) * resolution + origin)   subdivision.py:55-65
int cc_launch_make_blocks_subdiv(const int64_t *d_int_corners, uint32_t n, cc_level_geom g,
                                 cc_block_desc *d_blocks, void *stream);
// children int corners from hits: child = parent + (x,y,z) * int_step         subdivision.py:91-94
int cc_launch_expand_children(const int64_t *d_parent_corners, const uint32_t *d_hit_block,
                              const uint8_t *d_hit_xyz, uint32_t n_hits, int64_t int_step,
                              uint32_t rank, uint32_t world, int64_t *d_child_corners, void *stream);
// mass properties: float64 corner chain + per-block integrals
int cc_launch_mass_make_blocks(const double *d_corners, uint32_t n, double s, cc_block_desc *d_blocks,
                               void *stream);
int cc_launch_mass_expand_children(const double *d_parent_corners, const uint32_t *d_hit_block,
                                   const uint8_t *d_hit_xyz, uint32_t n_hits, double s,
                                   uint32_t rank, uint32_t world, double *d_child_corners, void *stream);
int cc_launch_mass_integrals(const double *d_corners, const uint32_t *d_sums, uint32_t n_blocks, double s,
                             double *d_integrals /* [10] accumulated with Kahan, single CTA */,
                             void *stream);

#endif
