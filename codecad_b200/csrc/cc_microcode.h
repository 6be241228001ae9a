// Device microcode: the decoded form of codecad's float32 node program.
//
// The wire format (one float word `opcode*512 + secondary_register` + parameters,
// /root/reference/codecad/nodes/program.py:55-71) is what Python hands us; it is
// decoded ONCE per program by cc_program.cpp into this layout, which is what the
// interpreter kernel walks:
//
//   * every instruction starts on a 16-byte boundary: word 0 is the header, words
//     1.. are fp32 parameters, length padded to a multiple of 4 words, so that a warp
//     fetches header + parameters with a few 128-bit broadcast loads;
//   * header = micro-op (8 bit) | src slot (9 bit) | dst slot (9 bit) | length/4 (6 bit);
//     the length lives in the header so that the interpreter advances its program counter
//     in ONE place (the loop latch) from a warp-uniform value — the shape of code for which
//     nvcc keeps decode and dispatch in the uniform datapath;
//     `_store` instructions of the wire format are folded into the producing
//     instruction's dst field and registers are renamed to a dense set of "slots" by a
//     liveness pass;
//   * parameter-only arithmetic is hoisted here (quaternion -> 3x3 matrix, polygon
//     edge tables, gear constants) — the canonical cc-arith definitions of DESIGN.md.
#ifndef CC_MICROCODE_H
#define CC_MICROCODE_H

#include <stdint.h>

enum cc_mop : uint32_t {
    MOP_RETURN = 0,
    MOP_LOAD,          // L = slot[src]
    MOP_NOP,           // only carries a dst (stand-alone _store)
    MOP_RECTANGLE,     // hw, hh
    MOP_CIRCLE,        // r
    MOP_REGPOLY,       // piOverN, r, r*sin, r*cos, 2*piOverN
    MOP_POLYGON,       // n, word offset of its edge table: n * (px, py, dx, dy, 1/|d|^2, cy) + group entries, stored after RETURN
    MOP_SPHERE,        // r
    MOP_HALF_SPACE,
    MOP_REV_TO,
    MOP_TWIST_TO,      // r, twist
    MOP_T_INIT,        // m[9], o[3]   (reads the kernel's grid point)
    MOP_T_TO,          // m[9], o[3]
    MOP_T_FROM,        // m[9] (already divided by |q|^2), scale
    MOP_MIRROR,
    MOP_SYM_TO,
    MOP_OFFSET,        // d
    MOP_SHELL,         // half thickness
    MOP_REPETITION,    // ox, oy, oz
    MOP_CREP_TO,       // piOverN, 2*piOverN
    MOP_CREP_FROM,     // piOverN, 2*piOverN            (src = point)
    MOP_GEAR,          // baseRadius, toothAngle, halfToothBaseAngle, 2*toothAngle, baseRadius^2
    MOP_EXTRUSION,     // halfH                          (src = point)
    MOP_REV_FROM,      //                                (src = point)
    MOP_TWIST_FROM,    // minorR, r, twist, min(1,lipschitz), padding   (src = point)
    MOP_SYM_FROM,      //                                (src = point)
    MOP_UNION,         // sharp: min                     (src = second operand)
    MOP_UNION_R,       // r >= 0
    MOP_ISECT,
    MOP_ISECT_R,
    MOP_SUB,
    MOP_SUB_R,
    MOP_PRIM_CIRCLE,   // fused T_INIT -> circle -> extrusion -> offset -> T_FROM   (28 words)
    MOP_PRIM_RECT,     // fused T_INIT -> rectangle -> extrusion -> offset -> T_FROM
    // "_M" variants: some matrix coefficient is exactly zero, and cc-arith omits such terms from
    // the row sums (cc_ops.cuh).  Same parameter layout plus one word of zero-masks (T_INIT/T_TO:
    // word 13, T_FROM: word 11, PRIM: word 27 = mask | mask_from << 9).  Separate micro-ops keep
    // the interpreter's code for full matrices (every primitive of a rotated scene) untouched.
    MOP_T_INIT_M,
    MOP_T_TO_M,
    MOP_T_FROM_M,
    MOP_PRIM_CIRCLE_M,
    MOP_PRIM_RECT_M,
    MOP_COUNT
};

#define CC_SLOT_NONE 0x1FFu
#define CC_MAX_SLOTS 0x1FEu

#define CC_HDR(op, src, dst, len_words) \
    ((uint32_t)(op) | ((uint32_t)(src) << 8) | ((uint32_t)(dst) << 17) | ((uint32_t)((len_words) / 4) << 26))
#define CC_HDR_OP(h) ((h) & 0xFFu)
#define CC_HDR_SRC(h) (((h) >> 8) & 0x1FFu)
#define CC_HDR_DST(h) (((h) >> 17) & 0x1FFu)
#define CC_HDR_LEN(h) (((h) >> 26) * 4u)

// instruction length in 32-bit words (header included), multiple of 4
#define CC_LEN_0 4   // up to 3 parameters
#define CC_LEN_7 8   // up to 7 parameters
#define CC_LEN_T 16  // 12 parameters (+3 spare)
#define CC_LEN_PRIM 28
#define CC_CONST_WORDS 16128  // microcode words that fit the 63 KB __constant__ window

#endif
