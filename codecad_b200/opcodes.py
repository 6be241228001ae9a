"""Opcode table of the node-program wire format.

The wire format is the reference's float32 instruction stream
(/root/reference/codecad/nodes/program.py:55-71): each instruction is one word
``opcode * 512 + secondary_register`` followed by its parameters.  Opcode numbers
are the enumeration order of ``Node.node_types``
(/root/reference/codecad/nodes/node.py:12-56); 512 is EVAL_REGISTER_COUNT
(/root/reference/codecad/nodes/__init__.py:6).
"""

REGISTER_COUNT = 512
VARIABLE = -1  # polygon2d: 1 + 2*n words, first word = vertex count

# name, number of parameter words, arity (number of value inputs)
OPS = [
    ("_return", 0, 1),
    ("_store", 0, 1),
    ("_load", 0, 1),
    ("rectangle", 2, 1),
    ("circle", 1, 1),
    ("regular_polygon2d", 2, 1),
    ("polygon2d", VARIABLE, 1),
    ("sphere", 1, 1),
    ("half_space", 0, 1),
    ("revolution_to", 0, 1),
    ("twist_revolution_to", 2, 1),
    ("initial_transformation_to", 7, 0),
    ("transformation_to", 7, 1),
    ("transformation_from", 4, 1),
    ("mirror", 0, 1),
    ("symmetrical_to", 0, 1),
    ("offset", 1, 1),
    ("shell", 1, 1),
    ("repetition", 3, 1),
    ("circular_repetition_to", 1, 1),
    ("circular_repetition_from", 1, 2),
    ("involute_gear", 2, 1),
    ("extrusion", 1, 2),
    ("revolution_from", 0, 2),
    ("twist_revolution_from", 3, 2),
    ("symmetrical_from", 0, 2),
    ("union", 1, 2),
    ("intersection", 1, 2),
    ("subtraction", 1, 2),
]

OPCODE = {name: i for i, (name, _, _) in enumerate(OPS)}


def disassemble(words):
    """Yield (pc, name, secondary_register, params) for a wire-format program."""
    pc = 0
    n = len(words)
    while pc < n:
        ins = int(words[pc])
        op, reg = divmod(ins, REGISTER_COUNT)
        name, nparams, _ = OPS[op]
        if nparams == VARIABLE:
            nparams = 1 + 2 * int(words[pc + 1])
        yield pc, name, reg, [float(w) for w in words[pc + 1 : pc + 1 + nparams]]
        pc += 1 + nparams
        if name == "_return":
            break


# ---- device microcode (csrc/cc_microcode.h) -------------------------------------------------
MICRO_OPS = [
    "RETURN", "LOAD", "NOP", "RECTANGLE", "CIRCLE", "REGPOLY", "POLYGON", "SPHERE", "HALF_SPACE",
    "REV_TO", "TWIST_TO", "T_INIT", "T_TO", "T_FROM", "MIRROR", "SYM_TO", "OFFSET", "SHELL",
    "REPETITION", "CREP_TO", "CREP_FROM", "GEAR", "EXTRUSION", "REV_FROM", "TWIST_FROM", "SYM_FROM",
    "UNION", "UNION_R", "ISECT", "ISECT_R", "SUB", "SUB_R", "PRIM_CIRCLE", "PRIM_RECT",
    # "_M": the same op with a matrix that has zero coefficients (omitted from the row sums)
    "T_INIT_M", "T_TO_M", "T_FROM_M", "PRIM_CIRCLE_M", "PRIM_RECT_M",
]
SLOT_NONE = 0x1FF


def disassemble_microcode(code):
    """Yield (pc, name, src, dst, length_words) for a decoded program.  Header layout
    (csrc/cc_microcode.h): op 8 bit | src slot 9 bit | dst slot 9 bit | length/4 6 bit."""
    import numpy as _np
    code = _np.asarray(code, dtype=_np.uint32)
    pc = 0
    while pc < len(code):
        h = int(code[pc])
        name = MICRO_OPS[h & 0xFF]
        src, dst, length = (h >> 8) & 0x1FF, (h >> 17) & 0x1FF, (h >> 26) * 4
        yield pc, name, src, dst, length
        pc += length
        if name == "RETURN":
            break
