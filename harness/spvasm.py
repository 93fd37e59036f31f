"""Minimal SPIR-V word emitter (test/bench infrastructure).

There is no glslang/spirv-tools in the image and the reference ships no shader binaries
(SURVEY.md §4), so test shaders are assembled by hand.  Only what the reference's front end accepts
(spirv_compile.cpp pass-3 handlers) is exposed.  Layout follows the SPIR-V
logical module order so the output is also valid for real tools.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Sequence, Tuple

import numpy as np

MAGIC = 0x07230203
VERSION_1_0 = 0x00010000


class Op:
    Nop = 0
    FUnordEqual = 181
    FUnordNotEqual = 183
    FUnordLessThan = 185
    FUnordGreaterThan = 187
    FUnordLessThanEqual = 189
    FUnordGreaterThanEqual = 191
    ImageSampleExplicitLod = 88
    BitFieldInsert = 201
    BitFieldSExtract = 202
    BitFieldUExtract = 203
    FRem = 140
    FMod = 141
    Any = 154
    All = 155
    BitReverse = 204
    BitCount = 205
    Undef = 1
    Source = 3
    Name = 5
    MemberName = 6
    ExtInstImport = 11
    ExtInst = 12
    MemoryModel = 14
    EntryPoint = 15
    ExecutionMode = 16
    Capability = 17
    TypeVoid = 19
    TypeBool = 20
    TypeInt = 21
    TypeFloat = 22
    TypeVector = 23
    TypeMatrix = 24
    TypeImage = 25
    TypeSampledImage = 27
    TypeArray = 28
    TypeStruct = 30
    TypePointer = 32
    TypeFunction = 33
    ConstantTrue = 41
    ConstantFalse = 42
    Constant = 43
    ConstantComposite = 44
    ConstantNull = 46
    Function = 54
    FunctionParameter = 55
    FunctionEnd = 56
    FunctionCall = 57
    Variable = 59
    Load = 61
    Store = 62
    AccessChain = 65
    Decorate = 71
    MemberDecorate = 72
    VectorExtractDynamic = 77
    VectorInsertDynamic = 78
    VectorShuffle = 79
    CompositeConstruct = 80
    CompositeExtract = 81
    CompositeInsert = 82
    CopyObject = 83
    Transpose = 84
    ImageSampleImplicitLod = 87
    ConvertFToU = 109
    ConvertFToS = 110
    ConvertSToF = 111
    ConvertUToF = 112
    SNegate = 126
    UDiv = 134
    SDiv = 135
    UMod = 137
    SRem = 138
    SMod = 139
    IsNan = 156
    IsInf = 157
    LogicalEqual = 164
    LogicalNotEqual = 165
    LogicalOr = 166
    LogicalAnd = 167
    LogicalNot = 168
    INotEqual = 171
    UGreaterThan = 172
    SGreaterThan = 173
    UGreaterThanEqual = 174
    SGreaterThanEqual = 175
    ULessThan = 176
    ULessThanEqual = 178
    SLessThanEqual = 179
    ShiftRightLogical = 194
    ShiftRightArithmetic = 195
    BitwiseOr = 197
    BitwiseXor = 198
    Not = 200
    Switch = 251
    Bitcast = 124
    FNegate = 127
    IAdd = 128
    ISub = 130
    FAdd = 129
    FSub = 131
    IMul = 132
    FMul = 133
    FDiv = 136
    VectorTimesScalar = 142
    MatrixTimesScalar = 143
    VectorTimesMatrix = 144
    MatrixTimesVector = 145
    MatrixTimesMatrix = 146
    Dot = 148
    Select = 169
    IEqual = 170
    SLessThan = 177
    FOrdEqual = 180
    FOrdNotEqual = 182
    FOrdLessThan = 184
    FOrdGreaterThan = 186
    FOrdLessThanEqual = 188
    FOrdGreaterThanEqual = 190
    ShiftLeftLogical = 196
    BitwiseAnd = 199
    DPdx = 207
    DPdy = 208
    Phi = 245
    LoopMerge = 246
    SelectionMerge = 247
    Label = 248
    Branch = 249
    BranchConditional = 250
    Kill = 252
    Return = 253
    ReturnValue = 254


class SC:
    UniformConstant = 0
    Input = 1
    Uniform = 2
    Output = 3
    Function = 7
    PushConstant = 9


class Dec:
    Block = 2
    ColMajor = 5
    ArrayStride = 6
    MatrixStride = 7
    BuiltIn = 11
    Flat = 14
    Location = 30
    Binding = 33
    DescriptorSet = 34
    Offset = 35


class BuiltIn:
    Position = 0
    PointSize = 1
    VertexId = 5
    InstanceId = 6
    VertexIndex = 42
    InstanceIndex = 43


class GLSL:
    Determinant = 33
    Asinh = 22
    Acosh = 23
    Atanh = 24
    Tan = 15
    Asin = 16
    Acos = 17
    Atan = 18
    Sinh = 19
    Cosh = 20
    Tanh = 21
    Atan2 = 25
    Exp = 27
    Log = 28
    Exp2 = 29
    Log2 = 30
    FindILsb = 73
    FindSMsb = 74
    FindUMsb = 75
    NMin = 79
    NMax = 80
    NClamp = 81
    RoundEven = 2
    Trunc = 3
    FAbs = 4
    SAbs = 5
    FSign = 6
    SSign = 7
    Floor = 8
    Ceil = 9
    Fract = 10
    Radians = 11
    Degrees = 12
    UMin = 38
    SMin = 39
    UMax = 41
    SMax = 42
    UClamp = 44
    SClamp = 45
    Step = 48
    SmoothStep = 49
    Fma = 50
    Distance = 67
    FaceForward = 70
    Refract = 72
    Sin = 13
    Cos = 14
    Pow = 26
    Sqrt = 31
    InverseSqrt = 32
    MatrixInverse = 34
    FMin = 37
    FMax = 40
    FClamp = 43
    FMix = 46
    Length = 66
    Cross = 68
    Normalize = 69
    Reflect = 71


VERTEX = 0
FRAGMENT = 4


def _str_words(s: str) -> List[int]:
    b = s.encode("utf-8") + b"\0"
    while len(b) % 4:
        b += b"\0"
    return list(struct.unpack("<%dI" % (len(b) // 4), b))


def _f32_bits(v: float) -> int:
    return struct.unpack("<I", struct.pack("<f", float(np.float32(v))))[0]


class Module:
    def __init__(self) -> None:
        self._next = 1
        self.sec: Dict[str, List[int]] = {k: [] for k in (
            "cap", "ext", "mem", "entry", "mode", "debug", "annot", "types", "funcs")}
        self._types: Dict[Tuple, int] = {}
        self._consts: Dict[Tuple, int] = {}
        self._glsl = None
        self._cur = "funcs"
        self._emit("cap", Op.Capability, 1)  # Shader
        self._emit("mem", Op.MemoryModel, 0, 1)  # Logical GLSL450

    # ---- plumbing
    def new_id(self) -> int:
        i = self._next
        self._next += 1
        return i

    def _emit(self, sec: str, opcode: int, *words: int) -> None:
        self.sec[sec].append(((len(words) + 1) << 16) | opcode)
        self.sec[sec].extend(int(w) & 0xFFFFFFFF for w in words)

    def words(self) -> np.ndarray:
        body: List[int] = []
        for k in ("cap", "ext", "mem", "entry", "mode", "debug", "annot", "types", "funcs"):
            body += self.sec[k]
        return np.array([MAGIC, VERSION_1_0, 0, self._next, 0] + body, dtype=np.uint32)

    def glsl(self) -> int:
        if self._glsl is None:
            self._glsl = self.new_id()
            self._emit("ext", Op.ExtInstImport, self._glsl, *_str_words("GLSL.std.450"))
        return self._glsl

    # ---- debug / annotations
    def name(self, target: int, s: str) -> None:
        self._emit("debug", Op.Name, target, *_str_words(s))

    def decorate(self, target: int, dec: int, *params: int) -> None:
        self._emit("annot", Op.Decorate, target, dec, *params)

    def member_decorate(self, target: int, member: int, dec: int, *params: int) -> None:
        self._emit("annot", Op.MemberDecorate, target, member, dec, *params)

    def entry_point(self, model: int, func: int, name: str, interface: Sequence[int] = ()) -> None:
        self._emit("entry", Op.EntryPoint, model, func, *_str_words(name), *interface)
        if model == FRAGMENT:
            self._emit("mode", Op.ExecutionMode, func, 7)  # OriginUpperLeft

    # ---- types (memoised)
    def _type(self, key: Tuple, opcode: int, *words: int) -> int:
        if key not in self._types:
            i = self.new_id()
            self._emit("types", opcode, i, *words)
            self._types[key] = i
        return self._types[key]

    def t_void(self) -> int:
        return self._type(("void",), Op.TypeVoid)

    def t_bool(self) -> int:
        return self._type(("bool",), Op.TypeBool)

    def t_int(self, signed: int = 1) -> int:
        return self._type(("int", signed), Op.TypeInt, 32, signed)

    def t_float(self) -> int:
        return self._type(("float",), Op.TypeFloat, 32)

    def t_vec(self, elem: int, n: int) -> int:
        return self._type(("vec", elem, n), Op.TypeVector, elem, n)

    def t_fvec(self, n: int) -> int:
        return self.t_vec(self.t_float(), n)

    def t_mat(self, n: int) -> int:
        return self._type(("mat", n), Op.TypeMatrix, self.t_fvec(n), n)

    def t_array(self, elem: int, length: int) -> int:
        return self._type(("arr", elem, length), Op.TypeArray, elem, self.const_u(length))

    def t_struct(self, *members: int, tag: str = "") -> int:
        # structs are never merged unless explicitly tagged the same
        key = ("struct", tag or self._next, members)
        return self._type(key, Op.TypeStruct, *members)

    def t_ptr(self, sc: int, t: int) -> int:
        return self._type(("ptr", sc, t), Op.TypePointer, sc, t)

    def t_func(self, ret: int, *params: int) -> int:
        return self._type(("func", ret, params), Op.TypeFunction, ret, *params)

    def t_image(self, dim: int = 1) -> int:
        # sampled float image, Dim 1 = 2D, 3 = Cube
        return self._type(("img", dim), Op.TypeImage, self.t_float(), dim, 0, 0, 0, 1, 0)

    def t_sampled_image(self, dim: int = 1) -> int:
        return self._type(("simg", dim), Op.TypeSampledImage, self.t_image(dim))

    # ---- constants
    def const_bool(self, v: bool) -> int:
        key = ("b", bool(v))
        if key not in self._consts:
            i = self.new_id()
            self._emit("types", Op.ConstantTrue if v else Op.ConstantFalse, self.t_bool(), i)
            self._consts[key] = i
        return self._consts[key]

    def const_null(self, t: int) -> int:
        key = ("null", t)
        if key not in self._consts:
            i = self.new_id()
            self._emit("types", Op.ConstantNull, t, i)
            self._consts[key] = i
        return self._consts[key]

    def undef(self, t: int) -> int:
        """module-level OpUndef"""
        i = self.new_id()
        self._emit("types", Op.Undef, t, i)
        return i

    def const_f(self, v: float) -> int:
        key = ("f", _f32_bits(v))
        if key not in self._consts:
            i = self.new_id()
            self._emit("types", Op.Constant, self.t_float(), i, _f32_bits(v))
            self._consts[key] = i
        return self._consts[key]

    def const_i(self, v: int) -> int:
        key = ("i", v & 0xFFFFFFFF)
        if key not in self._consts:
            i = self.new_id()
            self._emit("types", Op.Constant, self.t_int(1), i, v & 0xFFFFFFFF)
            self._consts[key] = i
        return self._consts[key]

    def const_u(self, v: int) -> int:
        key = ("u", v & 0xFFFFFFFF)
        if key not in self._consts:
            i = self.new_id()
            self._emit("types", Op.Constant, self.t_int(0), i, v & 0xFFFFFFFF)
            self._consts[key] = i
        return self._consts[key]

    def const_fvec(self, *vals: float) -> int:
        ids = tuple(self.const_f(v) for v in vals)
        key = ("cc", ids)
        if key not in self._consts:
            i = self.new_id()
            self._emit("types", Op.ConstantComposite, self.t_fvec(len(vals)), i, *ids)
            self._consts[key] = i
        return self._consts[key]

    # ---- globals
    def variable(self, sc: int, pointee: int, name: str = "") -> int:
        i = self.new_id()
        self._emit("types", Op.Variable, self.t_ptr(sc, pointee), i, sc)
        if name:
            self.name(i, name)
        return i

    def input(self, pointee: int, location: int, name: str = "") -> int:
        v = self.variable(SC.Input, pointee, name)
        self.decorate(v, Dec.Location, location)
        return v

    def output(self, pointee: int, location: int, name: str = "") -> int:
        v = self.variable(SC.Output, pointee, name)
        self.decorate(v, Dec.Location, location)
        return v

    def builtin_input(self, pointee: int, builtin: int, name: str = "") -> int:
        v = self.variable(SC.Input, pointee, name)
        self.decorate(v, Dec.BuiltIn, builtin)
        return v

    def per_vertex_out(self) -> int:
        """gl_PerVertex { vec4 gl_Position; float gl_PointSize; } as glslang emits it."""
        st = self.t_struct(self.t_fvec(4), self.t_float(), tag="gl_PerVertex")
        self.member_decorate(st, 0, Dec.BuiltIn, BuiltIn.Position)
        self.member_decorate(st, 1, Dec.BuiltIn, BuiltIn.PointSize)
        self.decorate(st, Dec.Block)
        return self.variable(SC.Output, st, "gl_out")

    def ubo(self, struct_t: int, set_: int, binding: int, name: str = "") -> int:
        self.decorate(struct_t, Dec.Block)
        v = self.variable(SC.Uniform, struct_t, name)
        self.decorate(v, Dec.DescriptorSet, set_)
        self.decorate(v, Dec.Binding, binding)
        return v

    def push_constants(self, struct_t: int, name: str = "") -> int:
        self.decorate(struct_t, Dec.Block)
        return self.variable(SC.PushConstant, struct_t, name)

    def sampler2d(self, set_: int, binding: int, name: str = "", dim: int = 1) -> int:
        v = self.variable(SC.UniformConstant, self.t_sampled_image(dim), name)
        self.decorate(v, Dec.DescriptorSet, set_)
        self.decorate(v, Dec.Binding, binding)
        return v

    # ---- function bodies
    def begin_function(self, ret_t: int, func_t: int, params: Sequence[int] = ()) -> Tuple[int, List[int]]:
        f = self.new_id()
        self._emit("funcs", Op.Function, ret_t, f, 0, func_t)
        pids = []
        for pt in params:
            p = self.new_id()
            self._emit("funcs", Op.FunctionParameter, pt, p)
            pids.append(p)
        return f, pids

    def end_function(self) -> None:
        self._emit("funcs", Op.FunctionEnd)

    def label(self, i: int | None = None) -> int:
        i = i if i is not None else self.new_id()
        self._emit("funcs", Op.Label, i)
        self.cur_label = i
        return i

    def raw(self, opcode: int, result_t: int, result_id: int, *operands: int) -> int:
        """Like inst(), with a result id the caller reserved earlier (OpPhi operands refer to later ids)."""
        self._emit("funcs", opcode, result_t, result_id, *operands)
        return result_id

    def inst(self, opcode: int, result_t: int, *operands: int) -> int:
        """Emit an instruction with (result type, result id) and return the result id."""
        r = self.new_id()
        self._emit("funcs", opcode, result_t, r, *operands)
        return r

    def stmt(self, opcode: int, *operands: int) -> None:
        self._emit("funcs", opcode, *operands)

    # convenience wrappers
    def local(self, pointee: int, init: int | None = None) -> int:
        r = self.new_id()
        w = [self.t_ptr(SC.Function, pointee), r, SC.Function]
        if init is not None:
            w.append(init)
        self._emit("funcs", Op.Variable, *w)
        return r

    def load(self, t: int, ptr: int) -> int:
        return self.inst(Op.Load, t, ptr)

    def store(self, ptr: int, val: int) -> None:
        self.stmt(Op.Store, ptr, val)

    def access(self, sc: int, t: int, base: int, *indices: int) -> int:
        return self.inst(Op.AccessChain, self.t_ptr(sc, t), base, *indices)

    def ext(self, t: int, inst: int, *args: int) -> int:
        return self.inst(Op.ExtInst, t, self.glsl(), inst, *args)

    def extract(self, t: int, composite: int, *idx: int) -> int:
        return self.inst(Op.CompositeExtract, t, composite, *idx)

    def construct(self, t: int, *parts: int) -> int:
        return self.inst(Op.CompositeConstruct, t, *parts)

    def shuffle(self, t: int, a: int, b: int, *comps: int) -> int:
        return self.inst(Op.VectorShuffle, t, a, b, *comps)

    def ret(self) -> None:
        self.stmt(Op.Return)
