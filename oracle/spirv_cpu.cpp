/*
 * oracle/spirv_cpu.cpp — TEST INFRASTRUCTURE ONLY (parity oracle; never on the product path).
 * See spirv_cpu.h for scope and parity status ("parity unpinned": LLVM 6.0.0 JIT not buildable).
 *
 * Structure mirrors CompileFunction (spirv_compile.cpp:645-2432):
 *   pass 1  (:750-961)   names, decorations, types, global variable types, constants
 *   pass 2  (:975-1016)  functions and labels
 *   pass 3  (:1037-1892) per-opcode semantics            -> Interp::exec()
 *   wrappers(:1894-2372) VS / FS entry marshalling        -> run_vertex() / run_fragment()
 * Compile with -O2 -ffp-contract=off and no -ffast-math / -march=native.
 */
#include "spirv_cpu.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <map>
#include <set>
#include <vector>

namespace vor
{
// Extended mode: a handful of opcodes the reference asserts on (Appendix B "Not supported"), with their
// SPIR-V semantics. Off by default: the oracle then rejects exactly what the reference rejects.
static bool g_extended = false;
void set_extended(bool on)
{
  g_extended = on;
}

namespace
{
// SPIR-V 1.1 enumerants used by the reference (values from the public SPIR-V specification).
enum : uint32_t
{
  kMagic = 0x07230203,
  kVersionMax = 0x00010100,
};
enum Op : uint16_t
{
  OpSource = 3, OpSourceExtension = 4, OpName = 5, OpMemberName = 6, OpExtInstImport = 11,
  OpExtInst = 12, OpMemoryModel = 14, OpEntryPoint = 15, OpExecutionMode = 16, OpCapability = 17,
  OpTypeVoid = 19, OpTypeBool = 20, OpTypeInt = 21, OpTypeFloat = 22, OpTypeVector = 23,
  OpTypeMatrix = 24, OpTypeImage = 25, OpTypeSampledImage = 27, OpTypeArray = 28, OpTypeStruct = 30,
  OpTypePointer = 32, OpTypeFunction = 33, OpConstant = 43, OpConstantComposite = 44,
  OpFunction = 54, OpFunctionParameter = 55, OpFunctionEnd = 56, OpFunctionCall = 57,
  OpVariable = 59, OpLoad = 61, OpStore = 62, OpAccessChain = 65, OpDecorate = 71,
  OpMemberDecorate = 72, OpVectorShuffle = 79, OpCompositeConstruct = 80, OpCompositeExtract = 81,
  OpTranspose = 84, OpImageSampleImplicitLod = 87, OpConvertSToF = 111, OpFNegate = 127,
  // outside the reference's subset: accepted only in extended mode
  OpConvertFToS = 110, OpBitcast = 124, OpISub = 130, OpSelect = 169, OpFOrdEqual = 180,
  OpFOrdNotEqual = 182, OpFOrdGreaterThanEqual = 190, OpPhi = 245, OpKill = 252,
  OpConvertFToU = 109, OpConvertUToF = 112, OpSNegate = 126, OpUDiv = 134, OpSDiv = 135, OpUMod = 137, OpSRem = 138,
  OpSMod = 139, OpIsNan = 156, OpIsInf = 157, OpLogicalEqual = 164, OpLogicalNotEqual = 165, OpLogicalOr = 166,
  OpLogicalAnd = 167, OpLogicalNot = 168, OpINotEqual = 171, OpUGreaterThan = 172, OpSGreaterThan = 173,
  OpUGreaterThanEqual = 174, OpSGreaterThanEqual = 175, OpULessThan = 176, OpULessThanEqual = 178,
  OpSLessThanEqual = 179, OpShiftRightLogical = 194, OpShiftRightArithmetic = 195, OpBitwiseOr = 197,
  OpBitwiseXor = 198, OpNot = 200, OpSwitch = 251,
  OpNop = 0, OpUndef = 1, OpConstantTrue = 41, OpConstantFalse = 42, OpConstantNull = 46, OpVectorExtractDynamic = 77,
  OpVectorInsertDynamic = 78, OpCompositeInsert = 82, OpCopyObject = 83,
  OpFRem = 140, OpFMod = 141, OpAny = 154, OpAll = 155, OpBitReverse = 204, OpBitCount = 205,
  OpImageSampleExplicitLod = 88, OpBitFieldInsert = 201, OpBitFieldSExtract = 202, OpBitFieldUExtract = 203,
  OpFUnordEqual = 181, OpFUnordNotEqual = 183, OpFUnordLessThan = 185, OpFUnordGreaterThan = 187,
  OpFUnordLessThanEqual = 189, OpFUnordGreaterThanEqual = 191,
  OpIAdd = 128, OpFAdd = 129, OpFSub = 131, OpIMul = 132, OpFMul = 133, OpFDiv = 136,
  OpVectorTimesScalar = 142, OpMatrixTimesScalar = 143, OpVectorTimesMatrix = 144,
  OpMatrixTimesVector = 145, OpMatrixTimesMatrix = 146, OpDot = 148, OpIEqual = 170,
  OpSLessThan = 177, OpFOrdLessThan = 184, OpFOrdGreaterThan = 186, OpFOrdLessThanEqual = 188,
  OpShiftLeftLogical = 196, OpBitwiseAnd = 199, OpDPdx = 207, OpDPdy = 208, OpLoopMerge = 246,
  OpSelectionMerge = 247, OpLabel = 248, OpBranch = 249, OpBranchConditional = 250,
  OpReturn = 253, OpReturnValue = 254,
};
enum : uint32_t
{
  SC_UniformConstant = 0, SC_Input = 1, SC_Uniform = 2, SC_Output = 3, SC_Function = 7,
  SC_PushConstant = 9,
  Dec_Block = 2, Dec_BuiltIn = 11, Dec_Location = 30, Dec_Binding = 33, Dec_DescriptorSet = 34,
  Dec_Offset = 35,
  BI_Position = 0, BI_PointSize = 1, BI_ClipDistance = 3, BI_CullDistance = 4, BI_VertexId = 5,
  BI_InstanceId = 6, BI_VertexIndex = 42, BI_InstanceIndex = 43,
  Dim_Cube = 3,
  EM_Vertex = 0, EM_Fragment = 4,
  G_Sin = 13, G_Cos = 14, G_Pow = 26, G_Sqrt = 31, G_InverseSqrt = 32, G_MatrixInverse = 34,
  G_FMin = 37, G_FMax = 40, G_FClamp = 43, G_FMix = 46, G_Length = 66, G_Cross = 68,
  G_Normalize = 69, G_Reflect = 71,
  // extended mode only
  G_RoundEven = 2, G_Trunc = 3, G_FAbs = 4, G_SAbs = 5, G_FSign = 6, G_SSign = 7, G_Floor = 8, G_Ceil = 9,
  G_Fract = 10, G_Radians = 11, G_Degrees = 12, G_UMin = 38, G_SMin = 39, G_UMax = 41, G_SMax = 42,
  G_UClamp = 44, G_SClamp = 45, G_Step = 48, G_SmoothStep = 49, G_Fma = 50, G_Distance = 67, G_FaceForward = 70,
  G_Refract = 72,
  G_FindILsb = 73, G_FindSMsb = 74, G_FindUMsb = 75, G_NMin = 79, G_NMax = 80, G_NClamp = 81, G_Determinant = 33,
  // extended mode, libm here and the special-function unit on the GPU (as Sin/Cos/Pow): 1-LSB colour bar
  G_Tan = 15, G_Asin = 16, G_Acos = 17, G_Atan = 18, G_Sinh = 19, G_Cosh = 20, G_Tanh = 21, G_Atan2 = 25,
  G_Exp = 27, G_Log = 28, G_Exp2 = 29, G_Log2 = 30, G_Asinh = 22, G_Acosh = 23, G_Atanh = 24,
};


enum TKind { T_NONE, T_VOID, T_BOOL, T_INT, T_FLOAT, T_VEC, T_MAT, T_ARR, T_STRUCT, T_PTR, T_FUNC, T_IMAGE };

// LLVM types as the reference creates them (:819-904), with the x86-64 DataLayout sizes the JIT
// applies to GEPs: vectors are aligned to their size rounded up to a power of two (<3 x float>
// occupies 16 bytes), arrays/matrices are arrays of those, structs are natural (non-packed).
// SPIR-V Offset/ArrayStride/MatrixStride decorations are NOT consulted by the reference.
struct Type
{
  TKind kind = T_NONE;
  uint32_t width = 0;      // scalar bit width
  uint32_t elem = 0;       // element / pointee / column type id
  uint32_t count = 0;      // vector comps, array length, matrix columns
  uint32_t storage = 0;    // pointer storage class
  std::vector<uint32_t> members;
  std::vector<uint32_t> offsets;
  uint32_t size = 0, align = 1;
};

struct Val
{
  union
  {
    float f[16];
    uint32_t u[16];
    int32_t i[16];
    uint64_t p;
  };
};

struct IDDecoration
{
  uint32_t id, dec, param, member;
};

struct ExternalBinding
{
  uint32_t storageClass;
  uint32_t var;    // variable id
  IDDecoration decoration;
};

struct Func
{
  uint32_t id = 0, retType = 0;
  std::vector<uint32_t> params;
  std::vector<const uint32_t *> insts;
  std::map<uint32_t, size_t> labels;    // label id -> index into insts
};

struct GlobalVar
{
  uint32_t id, ptrType, storage;
  bool block;
  uint32_t offset;    // into the Globals struct
};
}    // namespace

struct Entry
{
  const Module *mod;
  uint32_t model;
  uint32_t func;
  std::string name;
};

struct Module
{
  uint64_t serial = 0;    // unique per compile: a freed module's address may be reused by the next one
  std::vector<uint32_t> code;
  uint32_t idbound = 0;
  uint32_t glsl = 0;
  std::vector<Type> types;
  std::vector<uint32_t> valtype;    // result type of each value id
  std::vector<Val> consts;
  std::vector<uint8_t> isconst;
  std::vector<IDDecoration> decorations;
  std::set<uint32_t> blocks, cube;
  std::map<uint32_t, uint32_t> ptrtypes;
  std::vector<GlobalVar> globals;
  std::map<uint32_t, size_t> globalIndex;
  uint32_t globalsSize = 0;
  std::vector<ExternalBinding> externals;
  std::map<uint32_t, Func> funcs;
  std::vector<Entry> entries;
  std::map<uint32_t, uint32_t> descset;
};

namespace
{
struct CompileError
{
  std::string msg;
};
#define FAIL(...)                          \
  do                                       \
  {                                        \
    char _b[256];                          \
    snprintf(_b, sizeof(_b), __VA_ARGS__); \
    throw CompileError{_b};                \
  } while(0)

static uint32_t nextpow2(uint32_t v)
{
  uint32_t p = 1;
  while(p < v)
    p <<= 1;
  return p;
}

static uint32_t alignup(uint32_t v, uint32_t a)
{
  return (v + a - 1) / a * a;
}

static void layout(Module &m, uint32_t id)
{
  Type &t = m.types[id];
  switch(t.kind)
  {
    case T_BOOL: t.size = 1; t.align = 1; break;
    case T_INT:
    case T_FLOAT: t.size = t.width / 8; t.align = t.size; break;
    case T_VEC:
    {
      uint32_t raw = m.types[t.elem].size * t.count;
      t.align = nextpow2(raw);
      t.size = alignup(raw, t.align);
      break;
    }
    case T_MAT:
    case T_ARR:
      t.align = m.types[t.elem].align;
      t.size = m.types[t.elem].size * t.count;
      break;
    case T_STRUCT:
    {
      uint32_t off = 0, al = 1;
      for(uint32_t mid : t.members)
      {
        const Type &mt = m.types[mid];
        off = alignup(off, mt.align);
        t.offsets.push_back(off);
        off += mt.size;
        al = std::max(al, mt.align);
      }
      t.align = al;
      t.size = alignup(off, al);
      break;
    }
    case T_PTR:
    case T_IMAGE: t.size = 8; t.align = 8; break;
    default: t.size = 0; t.align = 1; break;
  }
}

// number of 32-bit lanes a value of this type has in a Val (vectors: comps; scalars 1)
static uint32_t comps(const Module &m, uint32_t tid)
{
  const Type &t = m.types[tid];
  return t.kind == T_VEC ? t.count : 1;
}

static bool isAggregate(const Module &m, uint32_t tid)
{
  TKind k = m.types[tid].kind;
  return k == T_MAT || k == T_ARR;
}

// Val <-> memory following the layout above. Matrices/arrays of vectors live in a Val with a
// fixed stride of 4 lanes per column (what the reference's alloca'd [N x <4 x float>] /
// [N x <3 x float>] temporaries have, :1344-1351).
static void loadVal(const Module &m, uint32_t tid, const uint8_t *p, Val &out)
{
  const Type &t = m.types[tid];
  switch(t.kind)
  {
    case T_BOOL: out.u[0] = *p ? 1 : 0; break;
    case T_INT:
    case T_FLOAT: memcpy(&out.u[0], p, 4); break;
    case T_VEC: memcpy(&out.u[0], p, 4 * t.count); break;
    case T_MAT:
    case T_ARR:
    {
      const Type &e = m.types[t.elem];
      for(uint32_t c = 0; c < t.count && c < 4; c++)
        memcpy(&out.u[c * 4], p + c * e.size, 4 * (e.kind == T_VEC ? e.count : 1));
      break;
    }
    case T_PTR:
    case T_IMAGE: memcpy(&out.p, p, 8); break;
    default: break;
  }
}

static void storeVal(const Module &m, uint32_t tid, uint8_t *p, const Val &in)
{
  const Type &t = m.types[tid];
  switch(t.kind)
  {
    case T_BOOL: *p = in.u[0] & 1; break;
    case T_INT:
    case T_FLOAT: memcpy(p, &in.u[0], 4); break;
    case T_VEC: memcpy(p, &in.u[0], 4 * t.count); break;
    case T_MAT:
    case T_ARR:
    {
      const Type &e = m.types[t.elem];
      for(uint32_t c = 0; c < t.count && c < 4; c++)
        memcpy(p + c * e.size, &in.u[c * 4], 4 * (e.kind == T_VEC ? e.count : 1));
      break;
    }
    case T_PTR:
    case T_IMAGE: memcpy(p, &in.p, 8); break;
    default: break;
  }
}

// the reference's sorted-insert (:787-811): std::lower_bound on id only, so decorations of one id
// end up in REVERSE declaration order.
static std::vector<IDDecoration>::iterator decoLower(std::vector<IDDecoration> &d, uint32_t id)
{
  return std::lower_bound(d.begin(), d.end(), id,
                          [](const IDDecoration &a, uint32_t b) { return a.id < b; });
}

static void parse(Module &m)
{
  const uint32_t *pCode = m.code.data();
  size_t codeSize = m.code.size();
  if(codeSize < 5)
    FAIL("module too short");
  if(pCode[0] != kMagic)
    FAIL("bad magic");    // :651
  if(pCode[1] > kVersionMax)
    FAIL("SPIR-V version > 1.1");    // :652
  if(pCode[4] != 0)
    FAIL("schema != 0");    // :657
  m.idbound = pCode[3];
  m.types.resize(m.idbound);
  m.valtype.assign(m.idbound, 0);
  m.consts.resize(m.idbound);
  m.isconst.assign(m.idbound, 0);

  const uint32_t *opStart = pCode + 5, *opEnd = pCode + codeSize;
  std::vector<const uint32_t *> entries;

  auto chk = [&](uint32_t id) {
    if(id >= m.idbound)
      FAIL("id %u out of bound", id);
    return id;
  };

  // ---- pass 1 (:750-961)
  for(pCode = opStart; pCode < opEnd;)
  {
    uint16_t wc = pCode[0] >> 16;
    uint16_t op = pCode[0] & 0xffff;
    if(wc == 0 || pCode + wc > opEnd)
      FAIL("bad word count");
    switch(op)
    {
      case OpExtInstImport:
        m.glsl = pCode[1];
        if(strcmp((const char *)(pCode + 2), "GLSL.std.450"))
          FAIL("unknown ext inst set");    // :776
        break;
      case OpEntryPoint: entries.push_back(pCode); break;
      case OpDecorate:
      {
        auto it = decoLower(m.decorations, pCode[1]);
        if(wc == 3)
        {
          m.decorations.insert(it, {pCode[1], pCode[2], 0, ~0U});
          if(pCode[2] == Dec_Block)
            m.blocks.insert(pCode[1]);
        }
        else
          m.decorations.insert(it, {pCode[1], pCode[2], pCode[3], ~0U});
        break;
      }
      case OpMemberDecorate:
      {
        auto it = decoLower(m.decorations, pCode[1]);
        if(wc == 4)
          m.decorations.insert(it, {pCode[1], pCode[3], 0, pCode[2]});
        else
          m.decorations.insert(it, {pCode[1], pCode[3], pCode[4], pCode[2]});
        break;
      }
      case OpTypeVoid: m.types[chk(pCode[1])].kind = T_VOID; break;
      case OpTypeFloat:
      {
        Type &t = m.types[chk(pCode[1])];
        t.kind = T_FLOAT;
        t.width = pCode[2];
        if(t.width != 32)
          FAIL("only 32-bit floats are exercised");
        layout(m, pCode[1]);
        break;
      }
      case OpTypeBool:
        m.types[chk(pCode[1])].kind = T_BOOL;
        layout(m, pCode[1]);
        break;
      case OpTypeInt:
      {
        Type &t = m.types[chk(pCode[1])];
        t.kind = T_INT;
        t.width = pCode[2];    // signedness ignored (:842)
        if(t.width != 32)
          FAIL("only 32-bit ints are exercised");
        layout(m, pCode[1]);
        break;
      }
      case OpTypeVector:
      {
        Type &t = m.types[chk(pCode[1])];
        t.kind = T_VEC;
        t.elem = chk(pCode[2]);
        t.count = pCode[3];
        if(t.count < 2 || t.count > 4)
          FAIL("vector size");
        layout(m, pCode[1]);
        break;
      }
      case OpTypeArray:
      {
        Type &t = m.types[chk(pCode[1])];
        t.kind = T_ARR;
        t.elem = chk(pCode[2]);
        if(!m.isconst[chk(pCode[3])])
          FAIL("array length not a constant");
        t.count = m.consts[pCode[3]].u[0];
        layout(m, pCode[1]);
        break;
      }
      case OpTypeMatrix:
      {
        Type &t = m.types[chk(pCode[1])];
        t.kind = T_MAT;    // "implement matrix as just array" (:858)
        t.elem = chk(pCode[2]);
        t.count = pCode[3];
        layout(m, pCode[1]);
        break;
      }
      case OpTypePointer:
      {
        Type &t = m.types[chk(pCode[1])];
        t.kind = T_PTR;
        t.storage = pCode[2];
        t.elem = chk(pCode[3]);
        layout(m, pCode[1]);
        if(m.blocks.count(pCode[3]) && (pCode[2] == SC_Uniform || pCode[2] == SC_PushConstant))
          m.blocks.insert(pCode[1]);    // :868-870
        m.ptrtypes[pCode[1]] = pCode[3];
        break;
      }
      case OpTypeStruct:
      {
        Type &t = m.types[chk(pCode[1])];
        t.kind = T_STRUCT;
        for(uint16_t i = 2; i < wc; i++)
          t.members.push_back(chk(pCode[i]));
        layout(m, pCode[1]);
        break;
      }
      case OpTypeFunction: m.types[chk(pCode[1])].kind = T_FUNC; break;
      case OpTypeImage:
      case OpTypeSampledImage:
      {
        if(op == OpTypeImage && pCode[3] == Dim_Cube)
          m.cube.insert(pCode[1]);
        else if(op == OpTypeSampledImage && m.cube.count(pCode[2]))
          m.cube.insert(pCode[1]);
        m.types[chk(pCode[1])].kind = T_IMAGE;    // t_VkImage (:902)
        layout(m, pCode[1]);
        break;
      }
      case OpVariable:
      {
        // global variable (:910-934)
        if(m.types[chk(pCode[1])].kind != T_PTR)
          FAIL("variable type is not a pointer");
        if(wc != 4)
          FAIL("global initialisers not handled");    // :932
        GlobalVar g;
        g.id = chk(pCode[2]);
        g.ptrType = pCode[1];
        g.storage = pCode[3];
        g.block = m.blocks.count(pCode[1]) != 0;
        if(g.block)
          m.blocks.insert(pCode[2]);
        g.offset = 0;
        m.globalIndex[g.id] = m.globals.size();
        m.globals.push_back(g);
        m.valtype[g.id] = pCode[1];
        break;
      }
      case OpConstant:
      {
        const Type &t = m.types[chk(pCode[1])];
        if(t.kind != T_FLOAT && t.kind != T_INT)
          FAIL("OpConstant of non-scalar");
        m.consts[chk(pCode[2])].u[0] = pCode[3];
        m.isconst[pCode[2]] = 1;
        m.valtype[pCode[2]] = pCode[1];
        break;
      }
      case OpConstantComposite:
      {
        const Type &t = m.types[chk(pCode[1])];
        if(t.kind != T_VEC)
          FAIL("OpConstantComposite: vectors only");    // :947
        for(uint16_t i = 3; i < wc && i < 3 + 4; i++)
          m.consts[chk(pCode[2])].u[i - 3] = m.consts[chk(pCode[i])].u[0];
        m.isconst[pCode[2]] = 1;
        m.valtype[pCode[2]] = pCode[1];
        break;
      }
      case OpConstantTrue: case OpConstantFalse: case OpConstantNull: case OpUndef:    // extended mode
      {
        if(!g_extended)
          FAIL("Unhandled SPIR-V opcode %u", op);    // :1888
        const Type &t = m.types[chk(pCode[1])];
        const TKind ek = t.kind == T_VEC ? m.types[t.elem].kind : t.kind;
        if((t.kind != T_VEC && t.kind != T_FLOAT && t.kind != T_INT && t.kind != T_BOOL) ||
           (ek != T_FLOAT && ek != T_INT && ek != T_BOOL) || ((op == OpConstantTrue || op == OpConstantFalse) && t.kind != T_BOOL))
          FAIL("constant %u: scalars and vectors of float, int or bool only", op);
        for(int c = 0; c < 4; c++)
          m.consts[chk(pCode[2])].u[c] = op == OpConstantTrue ? 1u : 0u;
        m.isconst[pCode[2]] = 1;
        m.valtype[pCode[2]] = pCode[1];
        break;
      }
      default: break;
    }
    if(op == OpFunction)
      break;    // :957
    pCode += wc;
  }

  // Globals struct (:963-964): non-block variables hold the pointee, block variables hold a pointer
  {
    uint32_t off = 0;
    for(GlobalVar &g : m.globals)
    {
      uint32_t sz, al;
      if(g.block)
      {
        sz = 8;
        al = 8;
      }
      else
      {
        const Type &pt = m.types[m.types[g.ptrType].elem];
        sz = pt.size;
        al = pt.align;
      }
      off = alignup(off, al);
      g.offset = off;
      off += sz;
    }
    m.globalsSize = alignup(off ? off : 16, 16);
  }

  // ---- pass 2 + 3 bookkeeping (:975-1016, :1097-1181)
  Func *cur = NULL;
  for(pCode = opStart; pCode < opEnd;)
  {
    uint16_t wc = pCode[0] >> 16;
    uint16_t op = pCode[0] & 0xffff;
    switch(op)
    {
      case OpFunction:
      {
        Func &f = m.funcs[chk(pCode[2])];
        f.id = pCode[2];
        f.retType = chk(pCode[1]);
        cur = &f;
        m.valtype[pCode[2]] = pCode[4];
        break;
      }
      case OpFunctionParameter:
        if(!cur)
          FAIL("parameter outside function");
        cur->params.push_back(chk(pCode[2]));
        m.valtype[pCode[2]] = chk(pCode[1]);
        break;
      case OpFunctionEnd: cur = NULL; break;
      case OpLabel:
        if(!cur)
          FAIL("label outside function");
        cur->labels[chk(pCode[1])] = cur->insts.size();
        cur->insts.push_back(pCode);
        break;
      case OpVariable:
        if(cur)
        {
          if(pCode[3] != SC_Function)
            FAIL("non-Function variable in function");    // :1101
          m.valtype[chk(pCode[2])] = chk(pCode[1]);
          cur->insts.push_back(pCode);
        }
        else
        {
          // externals (:1110-1130)
          uint32_t searchid = pCode[2];
          auto it = decoLower(m.decorations, searchid);
          if(it == m.decorations.end() || it->id != searchid)
          {
            searchid = m.ptrtypes[pCode[1]];
            it = decoLower(m.decorations, searchid);
          }
          if(pCode[3] <= SC_Output || pCode[3] == SC_PushConstant)
            for(; it != m.decorations.end() && it->id == searchid; ++it)
              m.externals.push_back({pCode[3], pCode[2], *it});
        }
        break;
      default:
        if(cur)
        {
          // every opcode the reference's pass 3 would reach inside a function body
          switch(op)
          {
            case OpFOrdLessThan: case OpFOrdLessThanEqual: case OpFOrdGreaterThan:
            case OpSLessThan: case OpIEqual: case OpShiftLeftLogical: case OpBitwiseAnd:
            case OpConvertSToF: case OpFunctionCall: case OpLoad: case OpAccessChain:
            case OpVectorTimesMatrix: case OpMatrixTimesVector: case OpMatrixTimesMatrix:
            case OpMatrixTimesScalar: case OpTranspose: case OpVectorTimesScalar: case OpFMul:
            case OpFDiv: case OpFAdd: case OpFSub: case OpFNegate: case OpIMul: case OpIAdd:
            case OpDPdx: case OpDPdy: case OpExtInst: case OpDot: case OpCompositeExtract:
            case OpCompositeConstruct: case OpVectorShuffle: case OpImageSampleImplicitLod:
              m.valtype[chk(pCode[2])] = chk(pCode[1]);
              cur->insts.push_back(pCode);
              break;
            case OpConvertFToS: case OpBitcast: case OpISub: case OpSelect: case OpFOrdEqual:
            case OpFOrdNotEqual: case OpFOrdGreaterThanEqual: case OpPhi:
            case OpConvertFToU: case OpConvertUToF: case OpSNegate: case OpUDiv: case OpSDiv: case OpUMod: case OpSRem:
            case OpSMod: case OpIsNan: case OpIsInf: case OpLogicalEqual: case OpLogicalNotEqual: case OpLogicalOr:
            case OpLogicalAnd: case OpLogicalNot: case OpINotEqual: case OpUGreaterThan: case OpSGreaterThan:
            case OpUGreaterThanEqual: case OpSGreaterThanEqual: case OpULessThan: case OpULessThanEqual:
            case OpSLessThanEqual: case OpShiftRightLogical: case OpShiftRightArithmetic: case OpBitwiseOr:
            case OpFRem: case OpFMod: case OpAny: case OpAll: case OpBitReverse: case OpBitCount:
            case OpImageSampleExplicitLod: case OpBitFieldInsert: case OpBitFieldSExtract: case OpBitFieldUExtract:
            case OpFUnordEqual: case OpFUnordNotEqual: case OpFUnordLessThan: case OpFUnordGreaterThan:
            case OpFUnordLessThanEqual: case OpFUnordGreaterThanEqual:
            case OpBitwiseXor: case OpNot: case OpUndef: case OpVectorExtractDynamic: case OpVectorInsertDynamic:
            case OpCompositeInsert: case OpCopyObject:
              if(!g_extended)
                FAIL("Unhandled SPIR-V opcode %u", op);    // :1888
              m.valtype[chk(pCode[2])] = chk(pCode[1]);
              cur->insts.push_back(pCode);
              break;
            case OpNop:
              if(!g_extended)
                FAIL("Unhandled SPIR-V opcode %u", op);    // :1888
              break;
            case OpSwitch:
            case OpKill:
              if(!g_extended)
                FAIL("Unhandled SPIR-V opcode %u", op);    // :1888
              cur->insts.push_back(pCode);
              break;
            case OpStore: case OpBranch: case OpBranchConditional: case OpReturn:
            case OpReturnValue:
              cur->insts.push_back(pCode);
              break;
            case OpSelectionMerge: case OpLoopMerge: break;    // :1245-1250
            // module-level opcodes that pass 3 skips (:1058-1091) may legally appear here too
            case OpName: case OpMemberName: case OpSource: case OpSourceExtension: break;
            default: FAIL("Unhandled SPIR-V opcode %u", op);    // :1888
          }
        }
        else
        {
          switch(op)
          {
            case OpCapability: case OpMemoryModel: case OpExecutionMode: case OpExtInstImport:
            case OpSource: case OpSourceExtension: case OpMemberName: case OpName: case OpEntryPoint:
            case OpConstantTrue: case OpConstantFalse: case OpConstantNull: case OpUndef:
            case OpDecorate: case OpMemberDecorate: case OpConstantComposite: case OpConstant:
            case OpTypeVoid: case OpTypeBool: case OpTypeInt: case OpTypeFloat: case OpTypeVector:
            case OpTypeArray: case OpTypeMatrix: case OpTypePointer: case OpTypeStruct:
            case OpTypeFunction: case OpTypeImage: case OpTypeSampledImage: break;
            default: FAIL("Unhandled SPIR-V opcode %u", op);    // :1888
          }
        }
        break;
    }
    pCode += wc;
  }

  // static cube propagation (:1290-1291): an OpLoad whose result TYPE is cube marks its value
  for(auto &kv : m.funcs)
    for(const uint32_t *w : kv.second.insts)
      if((w[0] & 0xffff) == OpLoad && m.cube.count(w[1]))
        m.cube.insert(w[2]);

  // descriptor-set cache (:1907-1910)
  for(const ExternalBinding &ext : m.externals)
    if(ext.decoration.dec == Dec_DescriptorSet)
      m.descset[ext.decoration.id] = ext.decoration.param;

  for(const uint32_t *e : entries)
  {
    Entry en;
    en.mod = &m;
    en.model = e[1];
    en.func = e[2];
    en.name = (const char *)&e[3];
    if(en.model != EM_Vertex && en.model != EM_Fragment)
      FAIL("Unsupported execution model");    // :2369
    if(!m.funcs.count(en.func))
      FAIL("entry point function missing");
    m.entries.push_back(en);
  }
}

// ---- helper intrinsics (:423-550), restated verbatim in order of operations -----------------

static void Float4x4TimesVec4(const float *fmat, const float *vec, float *out)
{
  for(int row = 0; row < 4; row++)
  {
    out[row] = 0.0f;
    for(int col = 0; col < 4; col++)
      out[row] += fmat[col * 4 + row] * vec[col];
  }
}
static void Float3x3TimesVec3(const float *fmat, const float *vec, float *out)
{
  for(int row = 0; row < 3; row++)
  {
    out[row] = 0.0f;
    for(int col = 0; col < 3; col++)
      out[row] += fmat[col * 4 + row] * vec[col];
  }
}
static void Vec4TimesFloat4x4(const float *fmat, const float *vec, float *out)
{
  for(int row = 0; row < 4; row++)
  {
    out[row] = 0.0f;
    for(int col = 0; col < 4; col++)
      out[row] += fmat[row * 4 + col] * vec[col];
  }
}
static void Vec3TimesFloat3x3(const float *fmat, const float *vec, float *out)
{
  for(int row = 0; row < 3; row++)
  {
    out[row] = 0.0f;
    for(int col = 0; col < 3; col++)
      out[row] += fmat[row * 4 + col] * vec[col];
  }
}
static void Float4x4TimesFloat4x4(const float *a, const float *b, float *out)
{
  for(size_t x = 0; x < 4; x++)
    for(size_t y = 0; y < 4; y++)
      out[x * 4 + y] = b[x * 4 + 0] * a[0 * 4 + y] + b[x * 4 + 1] * a[1 * 4 + y] +
                       b[x * 4 + 2] * a[2 * 4 + y] + b[x * 4 + 3] * a[3 * 4 + y];
}
static void Float4x4TimesFloat(const float *a, float b, float *out)
{
  for(size_t x = 0; x < 4; x++)
    for(size_t y = 0; y < 4; y++)
      out[x * 4 + y] = a[x * 4 + y] * b;
}
static void Float4x4Transpose(const float *in, float *out)
{
  for(size_t x = 0; x < 4; x++)
    for(size_t y = 0; y < 4; y++)
      out[x * 4 + y] = in[y * 4 + x];
}

// CreateDot (:629-643)
static float dotN(const float *a, const float *b, int n)
{
  float accum = a[0] * b[0];
  if(n > 1)
    accum = accum + a[1] * b[1];
  if(n > 2)
    accum = accum + a[2] * b[2];
  if(n > 3)
    accum = accum + a[3] * b[3];
  return accum;
}

struct State
{
  const Module *mod = NULL;
  uint64_t serial = 0;
  std::vector<Val> vals;
  std::vector<uint8_t> arena;
  size_t top = 0;
  std::vector<uint8_t> globals;
  bool killed = false;    // extended mode: the invocation executed OpKill

  uint8_t *alloc(uint32_t size, uint32_t align)
  {
    size_t a = (top + align - 1) / align * align;
    if(a + size > arena.size())
    {
      fprintf(stderr, "vor: shader arena exhausted\n");
      abort();
    }
    top = a + size;
    return arena.data() + a;
  }
};

static State &stateFor(const Module *m)
{
  // the reference's threaded path calls shaders from 8 threads (rasterizer.cpp:49-72)
  static thread_local std::vector<State *> cache;
  for(State *s : cache)
    if(s->mod == m && s->serial == m->serial)
      return *s;
  State *s = new State;
  s->mod = m;
  s->serial = m->serial;
  s->vals = m->consts;
  s->arena.resize(1 << 16);
  s->globals.assign(m->globalsSize, 0);
  for(const GlobalVar &g : m->globals)
    s->vals[g.id].p = (uint64_t)(uintptr_t)(s->globals.data() + g.offset);
  if(cache.size() >= 8)
  {
    delete cache.front();
    cache.erase(cache.begin());
  }
  cache.push_back(s);
  return *s;
}

struct Interp
{
  const Module &m;
  State &st;
  const ShaderEnv &env;

  uint32_t ncomp(uint32_t valueId) const { return comps(m, m.valtype[valueId]); }

  void exec(const Func &fn, const Val *args, Val *ret)
  {
    Val *V = st.vals.data();
    for(size_t i = 0; i < fn.params.size(); i++)
      V[fn.params[i]] = args[i];

    size_t pc = 0;
    const size_t n = fn.insts.size();
    uint32_t curLabel = 0, prevLabel = 0;    // for OpPhi (extended mode)
    while(pc < n)
    {
      const uint32_t *w = fn.insts[pc++];
      const uint16_t wc = w[0] >> 16;
      const uint16_t op = w[0] & 0xffff;
      switch(op)
      {
        case OpLabel:
          prevLabel = curLabel;
          curLabel = w[1];
          break;
        case OpPhi:    // extended mode: all phis of a block read their operands before any of them is written
        {
          size_t first = pc - 1, last = first;
          while(last + 1 < n && (fn.insts[last + 1][0] & 0xffff) == OpPhi)
            last++;
          std::vector<Val> fresh(last - first + 1);
          for(size_t i = first; i <= last; i++)
          {
            const uint32_t *ph = fn.insts[i];
            const uint16_t pwc = ph[0] >> 16;
            bool found = false;
            for(uint16_t j = 3; j + 1 < pwc; j += 2)
              if(ph[j + 1] == prevLabel)
              {
                fresh[i - first] = V[ph[j]];
                found = true;
              }
            if(!found)
            {
              fprintf(stderr, "vor: OpPhi without an operand for predecessor %u\n", prevLabel);
              abort();
            }
          }
          for(size_t i = first; i <= last; i++)
            V[fn.insts[i][2]] = fresh[i - first];
          pc = last + 1;
          break;
        }
        case OpKill:
          st.killed = true;
          return;
        case OpVariable:    // :1097-1107
        {
          const Type &pt = m.types[m.types[w[1]].elem];
          uint8_t *mem = st.alloc(pt.size ? pt.size : 4, pt.align ? pt.align : 4);
          V[w[2]].p = (uint64_t)(uintptr_t)mem;
          if(wc > 4)
            storeVal(m, m.types[w[1]].elem, mem, V[w[4]]);
          break;
        }
        // ---- logic / bitwise (:1187-1226)
        case OpFOrdLessThan:
          for(uint32_t c = 0, k = ncomp(w[3]); c < k; c++)
            V[w[2]].u[c] = V[w[3]].f[c] < V[w[4]].f[c];
          break;
        case OpFOrdLessThanEqual:
          for(uint32_t c = 0, k = ncomp(w[3]); c < k; c++)
            V[w[2]].u[c] = V[w[3]].f[c] <= V[w[4]].f[c];
          break;
        case OpFOrdGreaterThan:
          for(uint32_t c = 0, k = ncomp(w[3]); c < k; c++)
            V[w[2]].u[c] = V[w[3]].f[c] > V[w[4]].f[c];
          break;
        case OpSLessThan:
          for(uint32_t c = 0, k = ncomp(w[3]); c < k; c++)
            V[w[2]].u[c] = V[w[3]].i[c] < V[w[4]].i[c];
          break;
        case OpIEqual:
          for(uint32_t c = 0, k = ncomp(w[3]); c < k; c++)
            V[w[2]].u[c] = V[w[3]].u[c] == V[w[4]].u[c];
          break;
        case OpShiftLeftLogical:    // emits a logical shift RIGHT (:1214) — reference bug, kept
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
            V[w[2]].u[c] = V[w[3]].u[c] >> (V[w[4]].u[c] & 31);
          break;
        case OpBitwiseAnd:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
            V[w[2]].u[c] = V[w[3]].u[c] & V[w[4]].u[c];
          break;
        case OpConvertSToF:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
            V[w[2]].f[c] = (float)V[w[3]].i[c];
          break;
        // ---- extended mode (plain SPIR-V semantics; not in the reference)
        case OpFOrdGreaterThanEqual:
          for(uint32_t c = 0, k = ncomp(w[3]); c < k; c++)
            V[w[2]].u[c] = V[w[3]].f[c] >= V[w[4]].f[c];
          break;
        case OpFOrdEqual:
          for(uint32_t c = 0, k = ncomp(w[3]); c < k; c++)
            V[w[2]].u[c] = V[w[3]].f[c] == V[w[4]].f[c];
          break;
        case OpFOrdNotEqual:    // ordered: false when either operand is NaN
          for(uint32_t c = 0, k = ncomp(w[3]); c < k; c++)
            V[w[2]].u[c] = V[w[3]].f[c] < V[w[4]].f[c] || V[w[3]].f[c] > V[w[4]].f[c];
          break;
        // unordered comparisons: true when either operand is NaN, i.e. the negation of the opposite ordered one
#define VOR_UNORD(OP, ORDERED_OPPOSITE) \
        case OP: \
          for(uint32_t c = 0, k = ncomp(w[3]); c < k; c++) \
          { \
            const float a = V[w[3]].f[c], b = V[w[4]].f[c]; \
            V[w[2]].u[c] = !(ORDERED_OPPOSITE); \
          } \
          break;
        VOR_UNORD(OpFUnordEqual, a < b || a > b)
        VOR_UNORD(OpFUnordNotEqual, a == b)
        VOR_UNORD(OpFUnordLessThan, a >= b)
        VOR_UNORD(OpFUnordGreaterThan, a <= b)
        VOR_UNORD(OpFUnordLessThanEqual, a > b)
        VOR_UNORD(OpFUnordGreaterThanEqual, a < b)
#undef VOR_UNORD
        case OpSelect:    // component-wise; a scalar condition selects whole operands
        {
          const uint32_t k = comps(m, w[1]), kc = ncomp(w[3]);
          Val r;
          memset(&r, 0, sizeof(r));
          for(uint32_t c = 0; c < k; c++)
            r.u[c] = (V[w[3]].u[kc == 1 ? 0 : c] & 1) ? V[w[4]].u[c] : V[w[5]].u[c];
          V[w[2]] = r;
          break;
        }
        case OpISub:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
            V[w[2]].u[c] = V[w[3]].u[c] - V[w[4]].u[c];
          break;
        case OpBitcast:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
            V[w[2]].u[c] = V[w[3]].u[c];
          break;
        case OpConvertFToS:    // round toward zero; out-of-range and NaN give INT_MIN (cvttss2si)
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
          {
            const float f = V[w[3]].f[c];
            V[w[2]].i[c] = (f >= -2147483648.0f && f < 2147483648.0f) ? (int32_t)f : INT32_MIN;
          }
          break;
        // integers, logic, conversions (extended mode): explicit results where SPIR-V leaves them open —
        // x / 0 = 0, x % 0 = 0, INT_MIN / -1 = INT_MIN, INT_MIN % -1 = 0, shift counts modulo 32, float -> uint
        // outside [0, 2^32) = 0 — the same as the PTX back end emits
#define VOR_CMP(OP, FIELD, EXPR)                                  \
  case OP:                                                        \
    for(uint32_t c = 0, k = ncomp(w[3]); c < k; c++)              \
    {                                                             \
      const auto a = V[w[3]].FIELD[c], b = V[w[4]].FIELD[c];      \
      V[w[2]].u[c] = (EXPR) ? 1u : 0u;                            \
    }                                                             \
    break;
        VOR_CMP(OpINotEqual, u, a != b)
        VOR_CMP(OpUGreaterThan, u, a > b)
        VOR_CMP(OpSGreaterThan, i, a > b)
        VOR_CMP(OpUGreaterThanEqual, u, a >= b)
        VOR_CMP(OpSGreaterThanEqual, i, a >= b)
        VOR_CMP(OpULessThan, u, a < b)
        VOR_CMP(OpULessThanEqual, u, a <= b)
        VOR_CMP(OpSLessThanEqual, i, a <= b)
        VOR_CMP(OpLogicalEqual, u, (a & 1u) == (b & 1u))
        VOR_CMP(OpLogicalNotEqual, u, (a & 1u) != (b & 1u))
        VOR_CMP(OpLogicalOr, u, ((a | b) & 1u) != 0u)
        VOR_CMP(OpLogicalAnd, u, ((a & b) & 1u) != 0u)
#undef VOR_CMP
        case OpLogicalNot:
          for(uint32_t c = 0, k = ncomp(w[3]); c < k; c++)
            V[w[2]].u[c] = (V[w[3]].u[c] & 1u) ^ 1u;
          break;
        case OpIsNan:
          for(uint32_t c = 0, k = ncomp(w[3]); c < k; c++)
            V[w[2]].u[c] = V[w[3]].f[c] != V[w[3]].f[c];
          break;
        case OpIsInf:
          for(uint32_t c = 0, k = ncomp(w[3]); c < k; c++)
            V[w[2]].u[c] = (V[w[3]].u[c] & 0x7fffffffu) == 0x7f800000u;
          break;
        case OpSNegate:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
            V[w[2]].u[c] = 0u - V[w[3]].u[c];
          break;
        case OpNot:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
            V[w[2]].u[c] = ~V[w[3]].u[c];
          break;
        case OpBitwiseOr:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
            V[w[2]].u[c] = V[w[3]].u[c] | V[w[4]].u[c];
          break;
        case OpBitwiseXor:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
            V[w[2]].u[c] = V[w[3]].u[c] ^ V[w[4]].u[c];
          break;
        case OpShiftRightLogical:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
            V[w[2]].u[c] = V[w[3]].u[c] >> (V[w[4]].u[c] & 31);
          break;
        case OpShiftRightArithmetic:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
            V[w[2]].i[c] = V[w[3]].i[c] >> (V[w[4]].u[c] & 31);
          break;
        case OpUDiv:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
            V[w[2]].u[c] = V[w[4]].u[c] ? V[w[3]].u[c] / V[w[4]].u[c] : 0u;
          break;
        case OpUMod:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
            V[w[2]].u[c] = V[w[4]].u[c] ? V[w[3]].u[c] % V[w[4]].u[c] : 0u;
          break;
        case OpSDiv:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
          {
            const int32_t a = V[w[3]].i[c], b = V[w[4]].i[c];
            V[w[2]].i[c] = b == 0 ? 0 : (b == -1 ? (int32_t)(0u - (uint32_t)a) : a / b);
          }
          break;
        case OpSRem:
        case OpSMod:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
          {
            const int32_t a = V[w[3]].i[c], b = V[w[4]].i[c];
            int32_t r = (b == 0 || b == -1) ? 0 : a % b;
            if(op == OpSMod && r != 0 && ((r ^ b) < 0))
              r = (int32_t)((uint32_t)r + (uint32_t)b);
            V[w[2]].i[c] = r;
          }
          break;
        case OpConvertUToF:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
            V[w[2]].f[c] = (float)V[w[3]].u[c];
          break;
        case OpConvertFToU:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
          {
            const float f = V[w[3]].f[c];
            V[w[2]].u[c] = (f > -1.0f && f < 4294967296.0f) ? (uint32_t)(f < 0.0f ? 0.0f : f) : 0u;
          }
          break;
        case OpCopyObject: V[w[2]] = V[w[3]]; break;
        case OpFRem:    // x - y * trunc(x / y); OpFMod: x - y * floor(x / y); every step rounded on its own
        case OpFMod:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
          {
            const float x = V[w[3]].f[c], y = V[w[4]].f[c];
            const float q = x / y;
            const float t = op == OpFRem ? truncf(q) : floorf(q);
            const float pr = y * t;
            V[w[2]].f[c] = x - pr;
          }
          break;
        case OpAny:
        case OpAll:
        {
          uint32_t acc = op == OpAll ? 1u : 0u;
          for(uint32_t c = 0, k = ncomp(w[3]); c < k; c++)
            acc = op == OpAll ? (acc & V[w[3]].u[c] & 1u) : (acc | (V[w[3]].u[c] & 1u));
          V[w[2]].u[0] = acc;
          break;
        }
        case OpBitFieldSExtract:    // base, offset, count; offset + count <= 32 (anything else is undefined in SPIR-V)
        case OpBitFieldUExtract:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
          {
            const uint32_t off = V[w[4]].u[0] & 31u, cnt = V[w[5]].u[0];
            uint32_t x = 0;
            if(cnt)
            {
              const uint32_t mask = cnt >= 32u ? 0xffffffffu : (1u << cnt) - 1u;
              x = (V[w[3]].u[c] >> off) & mask;
              if(op == OpBitFieldSExtract && cnt < 32u && (x >> (cnt - 1u)) & 1u)
                x |= ~mask;
            }
            V[w[2]].u[c] = x;
          }
          break;
        case OpBitFieldInsert:    // base, insert, offset, count
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
          {
            const uint32_t off = V[w[5]].u[0] & 31u, cnt = V[w[6]].u[0];
            const uint32_t mask = (cnt >= 32u ? 0xffffffffu : (1u << cnt) - 1u) << off;
            V[w[2]].u[c] = (V[w[3]].u[c] & ~mask) | ((V[w[4]].u[c] << off) & mask);
          }
          break;
        case OpBitCount:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
            V[w[2]].u[c] = (uint32_t)__builtin_popcount(V[w[3]].u[c]);
          break;
        case OpBitReverse:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
          {
            uint32_t x = V[w[3]].u[c], r = 0;
            for(int b = 0; b < 32; b++, x >>= 1)
              r = (r << 1) | (x & 1u);
            V[w[2]].u[c] = r;
          }
          break;
        case OpUndef: memset(&V[w[2]], 0, sizeof(Val)); break;
        case OpCompositeInsert:    // object, composite, indexes
        {
          Val r = V[w[4]];
          if(isAggregate(m, m.valtype[w[4]]))
          {
            if(wc == 6)
              memcpy(&r.u[w[5] * 4], V[w[3]].u, 16);
            else
              r.u[w[5] * 4 + w[6]] = V[w[3]].u[0];
          }
          else
            r.u[w[5]] = V[w[3]].u[0];
          V[w[2]] = r;
          break;
        }
        case OpVectorExtractDynamic:    // an index past the end gives component 0
        {
          const uint32_t n = ncomp(w[3]), ix = V[w[4]].u[0];
          V[w[2]].u[0] = V[w[3]].u[ix < n ? ix : 0];
          break;
        }
        case OpVectorInsertDynamic:    // an index past the end changes nothing
        {
          Val r = V[w[3]];
          const uint32_t n = ncomp(w[3]), ix = V[w[5]].u[0];
          if(ix < n)
            r.u[ix] = V[w[4]].u[0];
          V[w[2]] = r;
          break;
        }
        case OpSwitch:
        {
          const uint32_t sel = V[w[1]].u[0];
          uint32_t target = w[2];
          for(uint16_t i = 3; i + 1 < wc; i += 2)
            if(w[i] == sel)
            {
              target = w[i + 1];
              break;
            }
          pc = fn.labels.at(target);
          break;
        }
        // ---- flow control (:1232-1279)
        case OpBranch: pc = fn.labels.at(w[1]); break;
        case OpBranchConditional: pc = fn.labels.at((V[w[1]].u[0] & 1) ? w[2] : w[3]); break;
        case OpFunctionCall:
        {
          const Func &callee = m.funcs.at(w[3]);
          Val a[16];
          for(uint16_t i = 4; i < wc && i < 20; i++)
            a[i - 4] = V[w[i]];
          Val r;
          memset(&r, 0, sizeof(r));
          exec(callee, a, &r);
          if(st.killed)
            return;
          V[w[2]] = r;
          break;
        }
        case OpReturn: return;
        case OpReturnValue:
          if(ret)
            *ret = V[w[1]];
          return;
        // ---- memory (:1285-1318)
        case OpLoad:
          loadVal(m, w[1], (const uint8_t *)(uintptr_t)V[w[3]].p, V[w[2]]);
          break;
        case OpStore:
        {
          uint32_t pointee = m.types[m.valtype[w[1]]].elem;
          storeVal(m, pointee, (uint8_t *)(uintptr_t)V[w[1]].p, V[w[2]]);
          break;
        }
        case OpAccessChain:
        {
          uint8_t *base = (uint8_t *)(uintptr_t)V[w[3]].p;
          if(m.blocks.count(w[3]))
            memcpy(&base, base, 8);    // load the buffer pointer out of the block variable (:1309-1313)
          uint32_t tid = m.types[m.valtype[w[3]]].elem;
          for(uint16_t i = 4; i < wc; i++)
          {
            const Type &t = m.types[tid];
            uint32_t idx = V[w[i]].u[0];
            if(t.kind == T_STRUCT)
            {
              base += t.offsets[idx];
              tid = t.members[idx];
            }
            else if(t.kind == T_ARR || t.kind == T_MAT)
            {
              base += idx * m.types[t.elem].size;
              tid = t.elem;
            }
            else if(t.kind == T_VEC)
            {
              base += idx * m.types[t.elem].size;
              tid = t.elem;
            }
          }
          V[w[2]].p = (uint64_t)(uintptr_t)base;
          break;
        }
        // ---- maths (:1324-1523)
        case OpVectorTimesMatrix:
        case OpMatrixTimesVector:
        {
          const Val &mat = (op == OpMatrixTimesVector) ? V[w[3]] : V[w[4]];
          const Val &vec = (op == OpMatrixTimesVector) ? V[w[4]] : V[w[3]];
          uint32_t vecsize = comps(m, w[1]);
          Val r;
          memset(&r, 0, sizeof(r));
          if(vecsize == 3)
            (op == OpVectorTimesMatrix ? Vec3TimesFloat3x3 : Float3x3TimesVec3)(mat.f, vec.f, r.f);
          else
            (op == OpVectorTimesMatrix ? Vec4TimesFloat4x4 : Float4x4TimesVec4)(mat.f, vec.f, r.f);
          V[w[2]] = r;
          break;
        }
        case OpMatrixTimesMatrix:
        {
          Val r;
          Float4x4TimesFloat4x4(V[w[3]].f, V[w[4]].f, r.f);
          V[w[2]] = r;
          break;
        }
        case OpMatrixTimesScalar:
        {
          Val r;
          Float4x4TimesFloat(V[w[3]].f, V[w[4]].f[0], r.f);
          V[w[2]] = r;
          break;
        }
        case OpTranspose:
        {
          Val r;
          Float4x4Transpose(V[w[3]].f, r.f);
          V[w[2]] = r;
          break;
        }
        case OpVectorTimesScalar:
        {
          float s = V[w[4]].f[0];
          for(uint32_t c = 0, k = ncomp(w[3]); c < k; c++)
            V[w[2]].f[c] = V[w[3]].f[c] * s;
          break;
        }
        case OpFMul:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
            V[w[2]].f[c] = V[w[3]].f[c] * V[w[4]].f[c];
          break;
        case OpFDiv:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
            V[w[2]].f[c] = V[w[3]].f[c] / V[w[4]].f[c];
          break;
        case OpFAdd:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
            V[w[2]].f[c] = V[w[3]].f[c] + V[w[4]].f[c];
          break;
        case OpFSub:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
            V[w[2]].f[c] = V[w[3]].f[c] - V[w[4]].f[c];
          break;
        case OpFNegate:    // IRBuilder::CreateFNeg in LLVM 6 = fsub -0.0, x
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
            V[w[2]].f[c] = -0.0f - V[w[3]].f[c];
          break;
        case OpIMul:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
            V[w[2]].u[c] = V[w[3]].u[c] * V[w[4]].u[c];
          break;
        case OpIAdd:
          for(uint32_t c = 0, k = comps(m, w[1]); c < k; c++)
            V[w[2]].u[c] = V[w[3]].u[c] + V[w[4]].u[c];
          break;
        case OpDPdx:    // placeholder shuffle {2,3,0} (:1515)
        {
          Val a = V[w[3]];
          V[w[2]].u[0] = a.u[2];
          V[w[2]].u[1] = a.u[3];
          V[w[2]].u[2] = a.u[0];
          break;
        }
        case OpDPdy:    // placeholder shuffle {1,3,2} (:1521)
        {
          Val a = V[w[3]];
          V[w[2]].u[0] = a.u[1];
          V[w[2]].u[1] = a.u[3];
          V[w[2]].u[2] = a.u[2];
          break;
        }
        case OpExtInst: extinst(w, wc); break;
        case OpDot: V[w[2]].f[0] = dotN(V[w[3]].f, V[w[4]].f, ncomp(w[3])); break;
        // ---- aggregates (:1755-1836)
        case OpCompositeExtract:
        {
          Val src = V[w[3]];
          if(isAggregate(m, m.valtype[w[3]]))
          {
            if(wc == 5)
              memcpy(V[w[2]].u, &src.u[w[4] * 4], 16);
            else if(wc == 6)
              V[w[2]].u[0] = src.u[w[4] * 4 + w[5]];
          }
          else
            V[w[2]].u[0] = src.u[w[4]];
          break;
        }
        case OpCompositeConstruct:
        {
          Val r;
          memset(&r, 0, sizeof(r));
          bool agg = isAggregate(m, w[1]);
          for(uint16_t i = 0; i < wc - 3; i++)
          {
            if(agg)
              memcpy(&r.u[i * 4], V[w[3 + i]].u, 16);
            else
              r.u[i] = V[w[3 + i]].u[0];    // one element per constituent (:1784)
          }
          V[w[2]] = r;
          break;
        }
        case OpVectorShuffle:
        {
          Val a = V[w[3]], b = V[w[4]];
          uint32_t na = ncomp(w[3]);
          Val r;
          memset(&r, 0, sizeof(r));
          for(uint16_t i = 5; i < wc; i++)
          {
            uint32_t idx = w[i];
            r.u[i - 5] = idx == 0xffffffffu ? 0 : (idx < na ? a.u[idx] : b.u[(idx - na) & 3]);
          }
          V[w[2]] = r;
          break;
        }
        // ---- texture (:1842-1886)
        case OpImageSampleExplicitLod:    // extended mode: the Lod operand is ignored (mip 0, like every sample)
        case OpImageSampleImplicitLod:
        {
          Val r;
          memset(&r, 0, sizeof(r));
          const void *img = (const void *)(uintptr_t)V[w[3]].p;
          const Val &c = V[w[4]];
          if(m.cube.count(w[3]))
            env.sample_cube(env.user, c.f[0], c.f[1], c.f[2], img, r.f);
          else
            env.sample_tex(env.user, c.f[0], c.f[1], img, 0, r.f);
          V[w[2]] = r;
          break;
        }
        default: fprintf(stderr, "vor: opcode %u reached exec\n", op); abort();
      }
    }
  }

  void extinst(const uint32_t *w, uint16_t wc)
  {
    Val *V = st.vals.data();
#define ARG(n) (V[w[5 + n]])
    const uint32_t k = comps(m, w[1]);
    Val r;
    memset(&r, 0, sizeof(r));
    switch(w[4])
    {
      case G_FMax:
      case G_FMin:    // select(olt/ogt(a,b), a, b) (:1535-1545)
        for(uint32_t c = 0; c < k; c++)
        {
          float a = ARG(0).f[c], b = ARG(1).f[c];
          bool sel = w[4] == G_FMin ? (a < b) : (a > b);
          r.f[c] = sel ? a : b;
        }
        break;
      case G_FClamp:    // :1546-1560
        for(uint32_t c = 0; c < k; c++)
        {
          float val = ARG(0).f[c], lo = ARG(1).f[c], hi = ARG(2).f[c];
          float upperClamped = (val < hi) ? val : hi;
          r.f[c] = (upperClamped > lo) ? upperClamped : lo;
        }
        break;
      case G_FMix:    // (1-a)*x + a*y (:1561-1579)
        for(uint32_t c = 0; c < k; c++)
        {
          float x = ARG(0).f[c], y = ARG(1).f[c], a = ARG(2).f[c];
          float xmul = 1.0f - a;
          r.f[c] = xmul * x + a * y;
        }
        break;
      case G_FAbs:    // extended mode
        for(uint32_t c = 0; c < k; c++)
          r.f[c] = fabsf(ARG(0).f[c]);
        break;
      case G_Floor:
        for(uint32_t c = 0; c < k; c++)
          r.f[c] = floorf(ARG(0).f[c]);
        break;
      case G_Fract:    // x - floor(x)
        for(uint32_t c = 0; c < k; c++)
          r.f[c] = ARG(0).f[c] - floorf(ARG(0).f[c]);
        break;
      // ---- the rest is extended mode only: GLSL.std.450 with its plain semantics
      case G_RoundEven:
        for(uint32_t c = 0; c < k; c++)
          r.f[c] = nearbyintf(ARG(0).f[c]);    // default rounding mode: to nearest even
        break;
      case G_Trunc:
        for(uint32_t c = 0; c < k; c++)
          r.f[c] = truncf(ARG(0).f[c]);
        break;
      case G_Ceil:
        for(uint32_t c = 0; c < k; c++)
          r.f[c] = ceilf(ARG(0).f[c]);
        break;
      case G_SAbs:
        for(uint32_t c = 0; c < k; c++)
          r.u[c] = ARG(0).i[c] < 0 ? 0u - ARG(0).u[c] : ARG(0).u[c];
        break;
      case G_FSign:
        for(uint32_t c = 0; c < k; c++)
          r.f[c] = ARG(0).f[c] > 0.0f ? 1.0f : (ARG(0).f[c] < 0.0f ? -1.0f : 0.0f);
        break;
      case G_SSign:
        for(uint32_t c = 0; c < k; c++)
          r.i[c] = ARG(0).i[c] > 0 ? 1 : (ARG(0).i[c] < 0 ? -1 : 0);
        break;
      case G_Radians:
        for(uint32_t c = 0; c < k; c++)
          r.f[c] = ARG(0).f[c] * 0.017453292519943295f;
        break;
      case G_Degrees:
        for(uint32_t c = 0; c < k; c++)
          r.f[c] = ARG(0).f[c] * 57.29577951308232f;
        break;
      case G_UMin:
        for(uint32_t c = 0; c < k; c++)
          r.u[c] = ARG(0).u[c] < ARG(1).u[c] ? ARG(0).u[c] : ARG(1).u[c];
        break;
      case G_SMin:
        for(uint32_t c = 0; c < k; c++)
          r.i[c] = ARG(0).i[c] < ARG(1).i[c] ? ARG(0).i[c] : ARG(1).i[c];
        break;
      case G_UMax:
        for(uint32_t c = 0; c < k; c++)
          r.u[c] = ARG(0).u[c] > ARG(1).u[c] ? ARG(0).u[c] : ARG(1).u[c];
        break;
      case G_SMax:
        for(uint32_t c = 0; c < k; c++)
          r.i[c] = ARG(0).i[c] > ARG(1).i[c] ? ARG(0).i[c] : ARG(1).i[c];
        break;
      case G_UClamp:    // min(max(x, lo), hi)
        for(uint32_t c = 0; c < k; c++)
        {
          uint32_t v = ARG(0).u[c] > ARG(1).u[c] ? ARG(0).u[c] : ARG(1).u[c];
          r.u[c] = v < ARG(2).u[c] ? v : ARG(2).u[c];
        }
        break;
      case G_SClamp:
        for(uint32_t c = 0; c < k; c++)
        {
          int32_t v = ARG(0).i[c] > ARG(1).i[c] ? ARG(0).i[c] : ARG(1).i[c];
          r.i[c] = v < ARG(2).i[c] ? v : ARG(2).i[c];
        }
        break;
      case G_Step:    // x < edge ? 0 : 1
        for(uint32_t c = 0; c < k; c++)
          r.f[c] = ARG(1).f[c] < ARG(0).f[c] ? 0.0f : 1.0f;
        break;
      case G_SmoothStep:
        for(uint32_t c = 0; c < k; c++)
        {
          float num = ARG(2).f[c] - ARG(0).f[c], den = ARG(1).f[c] - ARG(0).f[c];
          float q = num / den;
          float u = (q < 1.0f) ? q : 1.0f;
          float t = (u > 0.0f) ? u : 0.0f;
          float tt = t * t, two = 2.0f * t, rest = 3.0f - two;
          r.f[c] = tt * rest;
        }
        break;
      case G_Fma:
        for(uint32_t c = 0; c < k; c++)
          r.f[c] = fmaf(ARG(0).f[c], ARG(1).f[c], ARG(2).f[c]);
        break;
      case G_Distance:
      {
        uint32_t n = comps(m, m.valtype[w[5]]);
        float d[4] = {0, 0, 0, 0};
        for(uint32_t c = 0; c < n; c++)
          d[c] = ARG(0).f[c] - ARG(1).f[c];
        r.f[0] = sqrtf(dotN(d, d, n));
        break;
      }
      case G_FaceForward:    // dot(Nref, I) < 0 ? N : -N
      {
        uint32_t n = comps(m, m.valtype[w[5]]);
        float dd = dotN(ARG(2).f, ARG(1).f, n);
        for(uint32_t c = 0; c < n; c++)
          r.f[c] = dd < 0.0f ? ARG(0).f[c] : -0.0f - ARG(0).f[c];
        break;
      }
      case G_Refract:    // k = 1 - eta*eta*(1 - d*d), d = dot(N, I); k < 0 ? 0 : eta*I - (eta*d + sqrt(k))*N
      {
        uint32_t n = comps(m, m.valtype[w[5]]);
        float eta = ARG(2).f[0];
        float d = dotN(ARG(1).f, ARG(0).f, n);
        float dd = d * d, om = 1.0f - dd, ee = eta * eta, eo = ee * om, kk = 1.0f - eo;
        float ed = eta * d, sq = sqrtf(kk), t = ed + sq;
        for(uint32_t c = 0; c < n; c++)
        {
          float ei = eta * ARG(0).f[c], tn = t * ARG(1).f[c];
          r.f[c] = kk < 0.0f ? 0.0f : ei - tn;
        }
        break;
      }
      case G_FindILsb:    // index of the lowest set bit, -1 for 0
        for(uint32_t c = 0; c < k; c++)
          r.i[c] = ARG(0).u[c] ? __builtin_ctz(ARG(0).u[c]) : -1;
        break;
      case G_FindSMsb:    // highest bit that differs from the sign; -1 for 0 and -1
        for(uint32_t c = 0; c < k; c++)
        {
          const uint32_t x = ARG(0).i[c] < 0 ? ~ARG(0).u[c] : ARG(0).u[c];
          r.i[c] = x ? 31 - __builtin_clz(x) : -1;
        }
        break;
      case G_FindUMsb:
        for(uint32_t c = 0; c < k; c++)
          r.i[c] = ARG(0).u[c] ? 31 - __builtin_clz(ARG(0).u[c]) : -1;
        break;
      case G_NMin: case G_NMax: case G_NClamp:    // a NaN operand yields the other one
      {
        auto nsel = [](bool isMin, float x, float y) {
          if(x != x)
            return y;
          if(y != y)
            return x;
          return (isMin ? y < x : y > x) ? y : x;
        };
        for(uint32_t c = 0; c < k; c++)
          r.f[c] = w[4] == G_NClamp ? nsel(true, nsel(false, ARG(0).f[c], ARG(1).f[c]), ARG(2).f[c])
                                    : nsel(w[4] == G_NMin, ARG(0).f[c], ARG(1).f[c]);
        break;
      }
      case G_Determinant:    // cofactors along row 0; every product and sum rounded on its own, left to right
      {
        const Type &mt = m.types[m.valtype[w[5]]];
        const uint32_t n = mt.count;
        const float *a = ARG(0).f;    // a[col * 4 + row] for 3x3 and 4x4 alike
        auto e = [&](uint32_t rr, uint32_t cc) { return a[cc * 4 + rr]; };
        auto det3 = [&](const uint32_t *rw, const uint32_t *cl) {
          float p, q2;
          p = e(rw[1], cl[1]) * e(rw[2], cl[2]); q2 = e(rw[2], cl[1]) * e(rw[1], cl[2]);
          const float m0 = p - q2;
          p = e(rw[1], cl[0]) * e(rw[2], cl[2]); q2 = e(rw[2], cl[0]) * e(rw[1], cl[2]);
          const float m1 = p - q2;
          p = e(rw[1], cl[0]) * e(rw[2], cl[1]); q2 = e(rw[2], cl[0]) * e(rw[1], cl[1]);
          const float m2 = p - q2;
          float d = e(rw[0], cl[0]) * m0;
          p = e(rw[0], cl[1]) * m1;
          d = d - p;
          p = e(rw[0], cl[2]) * m2;
          return d + p;
        };
        if(n == 3)
        {
          const uint32_t rw[3] = {0, 1, 2}, cl[3] = {0, 1, 2};
          r.f[0] = det3(rw, cl);
        }
        else
        {
          const uint32_t rw[3] = {1, 2, 3};
          float d = 0.0f;
          for(uint32_t j = 0; j < 4; j++)
          {
            uint32_t cl[3], q = 0;
            for(uint32_t cc = 0; cc < 4; cc++)
              if(cc != j)
                cl[q++] = cc;
            const float t = e(0, j) * det3(rw, cl);
            d = j == 0 ? t : (j & 1u) ? d - t : d + t;
          }
          r.f[0] = d;
        }
        break;
      }
      // transcendental functions: libm (the GPU uses its special-function unit; neither is the reference's CRT)
#define VOR_LIBM1(G, fn) \
      case G: \
        for(uint32_t c = 0; c < k; c++) \
          r.f[c] = fn(ARG(0).f[c]); \
        break;
      VOR_LIBM1(G_Exp, expf) VOR_LIBM1(G_Exp2, exp2f) VOR_LIBM1(G_Log, logf) VOR_LIBM1(G_Log2, log2f)
      VOR_LIBM1(G_Tan, tanf) VOR_LIBM1(G_Sinh, sinhf) VOR_LIBM1(G_Cosh, coshf) VOR_LIBM1(G_Tanh, tanhf)
      VOR_LIBM1(G_Atan, atanf) VOR_LIBM1(G_Asin, asinf) VOR_LIBM1(G_Acos, acosf)
      VOR_LIBM1(G_Asinh, asinhf) VOR_LIBM1(G_Acosh, acoshf) VOR_LIBM1(G_Atanh, atanhf)
#undef VOR_LIBM1
      case G_Atan2:
        for(uint32_t c = 0; c < k; c++)
          r.f[c] = atan2f(ARG(0).f[c], ARG(1).f[c]);
        break;
      case G_Cos: r.f[0] = cosf(ARG(0).f[0]); break;      // llvm.cos.f32 -> CRT (not reproducible)
      case G_Sin: r.f[0] = sinf(ARG(0).f[0]); break;      // llvm.sin.f32 -> CRT (not reproducible)
      case G_Sqrt: r.f[0] = sqrtf(ARG(0).f[0]); break;    // llvm.sqrt.f32 -> sqrtss (exact)
      case G_InverseSqrt:    // scalar only; the vector form crashes in the reference (:1619)
        r.f[0] = 1.0f / sqrtf(ARG(0).f[0]);
        break;
      case G_Normalize:    // a * splat(1.0 / sqrt(dot(a,a))) (:1635-1647)
      {
        uint32_t n = comps(m, m.valtype[w[5]]);
        float len = sqrtf(dotN(ARG(0).f, ARG(0).f, n));
        float invlen = 1.0f / len;
        for(uint32_t c = 0; c < n; c++)
          r.f[c] = ARG(0).f[c] * invlen;
        break;
      }
      case G_Length:
        r.f[0] = sqrtf(dotN(ARG(0).f, ARG(0).f, comps(m, m.valtype[w[5]])));
        break;
      case G_Cross: r = ARG(0); break;    // "TODO": returns operand 0 (:1657-1663) — kept
      case G_Pow:
        for(uint32_t c = 0; c < k; c++)
          r.f[c] = powf(ARG(0).f[c], ARG(1).f[c]);    // llvm.pow.f32 -> CRT (not reproducible)
        break;
      case G_Reflect:    // I - (dot(I,N)*2)*N (:1690-1702)
      {
        uint32_t n = comps(m, m.valtype[w[5]]);
        float NdotI = dotN(ARG(0).f, ARG(1).f, n);
        float NdotI2 = NdotI * 2.0f;
        for(uint32_t c = 0; c < n; c++)
          r.f[c] = ARG(0).f[c] - NdotI2 * ARG(1).f[c];
        break;
      }
      case G_MatrixInverse:    // calls Float4x4Transpose, not inverse (:1721) — kept
        Float4x4Transpose(ARG(0).f, r.f);
        break;
      default: fprintf(stderr, "vor: ext inst %u reached exec\n", w[4]); abort();
    }
    V[w[2]] = r;
#undef ARG
  }
};

static void validateExt(const Module &m)
{
  for(auto &kv : m.funcs)
    for(const uint32_t *w : kv.second.insts)
    {
      uint16_t op = w[0] & 0xffff, wc = w[0] >> 16;
      if(op == OpExtInst)
      {
        if(w[3] != m.glsl)
          FAIL("ext inst set");
        switch(w[4])
        {
          case G_FMax: case G_FMin: case G_FClamp: case G_FMix: case G_Cos: case G_Sin:
          case G_Sqrt: case G_Normalize: case G_Length: case G_Cross: case G_Pow: case G_Reflect:
          case G_MatrixInverse: break;
          case G_FAbs: case G_Floor: case G_Fract: case G_RoundEven: case G_Trunc: case G_Ceil: case G_SAbs:
          case G_FSign: case G_SSign: case G_Radians: case G_Degrees: case G_UMin: case G_SMin: case G_UMax:
          case G_SMax: case G_UClamp: case G_SClamp: case G_Step: case G_SmoothStep: case G_Fma: case G_Distance:
          case G_FaceForward: case G_Refract: case G_FindILsb: case G_FindSMsb: case G_FindUMsb: case G_NMin:
          case G_NMax: case G_NClamp: case G_Tan: case G_Asin: case G_Acos: case G_Atan: case G_Sinh: case G_Cosh:
          case G_Tanh: case G_Atan2: case G_Exp: case G_Log: case G_Exp2: case G_Log2: case G_Asinh: case G_Acosh:
          case G_Atanh:
            if(!g_extended)
              FAIL("Unhandled GLSL extended instruction %u", w[4]);    // :1734
            break;
          case G_Determinant:
          {
            if(!g_extended)
              FAIL("Unhandled GLSL extended instruction %u", w[4]);    // :1734
            const Type &mt = m.types[m.valtype[w[5]]];
            if(mt.kind != T_MAT || (mt.count != 3 && mt.count != 4))
              FAIL("Determinant of 3x3 and 4x4 matrices only");
            break;
          }
          case G_InverseSqrt:
            if(m.types[w[1]].kind == T_VEC)
              FAIL("vector InverseSqrt crashes the reference (:1619)");
            break;
          default: FAIL("Unhandled GLSL extended instruction %u", w[4]);    // :1734
        }
      }
      else if(op == OpVectorTimesMatrix || op == OpMatrixTimesVector)
      {
        uint32_t vs = comps(m, w[1]);
        const Type &mt = m.types[m.valtype[op == OpMatrixTimesVector ? w[3] : w[4]]];
        if((vs != 3 && vs != 4) || mt.count != vs || m.types[mt.elem].count != vs)
          FAIL("only square 3/4 matrix multiplies");    // :1356-1357
      }
      else if(op == OpFunctionCall)
      {
        if(!m.funcs.count(w[3]))
          FAIL("call to unknown function");
        if(wc - 4 > 16)
          FAIL("too many call arguments");
      }
      else if(op == OpLoad || op == OpCompositeConstruct)
      {
        const Type &t = m.types[w[1]];
        if(t.kind == T_STRUCT || ((t.kind == T_ARR || t.kind == T_MAT) && t.count > 4))
          FAIL("value type too large for the oracle interpreter");
      }
      else if(op == OpBranch)
      {
        if(!kv.second.labels.count(w[1]))
          FAIL("branch to unknown label");
      }
      else if(op == OpBranchConditional)
      {
        if(!kv.second.labels.count(w[2]) || !kv.second.labels.count(w[3]))
          FAIL("branch to unknown label");
      }
      else if(op == OpSwitch)
      {
        if(wc < 3 || ((wc - 3) & 1) || !kv.second.labels.count(w[2]))
          FAIL("malformed OpSwitch");
        for(uint16_t i = 3; i + 1 < wc; i += 2)
          if(!kv.second.labels.count(w[i + 1]))
            FAIL("branch to unknown label");
      }
    }
}

static uint8_t *globalPtr(State &st, const Module &m, uint32_t var)
{
  return st.globals.data() + m.globals[m.globalIndex.at(var)].offset;
}
static uint32_t globalPointee(const Module &m, uint32_t var)
{
  return m.types[m.globals[m.globalIndex.at(var)].ptrType].elem;
}
}    // namespace

bool fetch_vertex_attr(uint32_t format, const uint8_t *ptr, float out[4])
{
  // GetVertexAttributeData (:572-627); VkFormat values from vulkan.h v42
  out[0] = 0;
  out[1] = 0;
  out[2] = 0;
  out[3] = 1;
  float f32[4];
  uint32_t u32;
  switch(format)
  {
    case 109: case 107: case 108:    // R32G32B32A32_SFLOAT / _UINT / _SINT
      memcpy(f32, ptr, 16);
      out[3] = f32[3]; out[2] = f32[2]; out[1] = f32[1]; out[0] = f32[0];
      return true;
    case 106: case 104: case 105:    // R32G32B32
      memcpy(f32, ptr, 12);
      out[2] = f32[2]; out[1] = f32[1]; out[0] = f32[0];
      return true;
    case 103: case 101: case 102:    // R32G32
      memcpy(f32, ptr, 8);
      out[1] = f32[1]; out[0] = f32[0];
      return true;
    case 100: case 98: case 99:    // R32
      memcpy(f32, ptr, 4);
      out[0] = f32[0];
      return true;
    case 37:    // R8G8B8A8_UNORM
      memcpy(&u32, ptr, 4);
      out[0] = float((u32 & 0x000000ff) >> 0x00) / 255.0f;
      out[1] = float((u32 & 0x0000ff00) >> 0x08) / 255.0f;
      out[2] = float((u32 & 0x00ff0000) >> 0x10) / 255.0f;
      out[3] = float((u32 & 0xff000000) >> 0x18) / 255.0f;
      return true;
    default: return false;    // assert(false && "Unhandled vertex attribute format")
  }
}

Module *compile(const uint32_t *code, size_t words, std::string *err)
{
  static std::atomic<uint64_t> nextSerial{1};
  Module *m = new Module;
  m->serial = nextSerial++;
  m->code.assign(code, code + words);
  try
  {
    parse(*m);
    validateExt(*m);
  }
  catch(const CompileError &e)
  {
    if(err)
      *err = e.msg;
    delete m;
    return NULL;
  }
  return m;
}

const Entry *find_entry(const Module *m, const char *name)
{
  for(const Entry &e : m->entries)
    if(e.name == name)
      return &e;
  return NULL;
}

void destroy(Module *m)
{
  delete m;
}

int entry_stage(const Entry *e)
{
  return (int)e->model;
}

// inputs common to both wrappers: UBO pointers, images, push constants
static void bindResources(const Module &m, State &st, const ShaderEnv &env, const ExternalBinding &ext,
                          bool fragment)
{
  uint32_t id = ext.decoration.id;
  uint8_t *val = globalPtr(st, m, ext.var);
  if(ext.decoration.dec == Dec_Binding)
  {
    uint32_t set = 0;
    auto it = m.descset.find(id);
    if(it != m.descset.end())
      set = it->second;
    if(m.blocks.count(id))
    {
      const uint8_t *p = env.buffer_ptr(env.user, set, ext.decoration.param);    // :2004-2012
      memcpy(val, &p, 8);
    }
    else if(fragment)
    {
      const void *img = env.image(env.user, set, ext.decoration.param);    // :2302-2309
      memcpy(val, &img, 8);
    }
    else
    {
      fprintf(stderr, "vor: image binding in a vertex shader (assert :2002)\n");
      abort();
    }
  }
  else if(ext.decoration.dec == Dec_Offset && ext.storageClass == SC_PushConstant)
  {
    const uint8_t *p = env.push_ptr(env.user, ext.decoration.param);    // :2017-2028
    memcpy(val, &p, 8);
  }
}

void run_vertex(const Entry *e, const ShaderEnv &env, uint32_t vertexIndex, float out[kVertexFloats])
{
  const Module &m = *e->mod;
  State &st = stateFor(&m);
  st.top = 0;

  // inputs (:1930-2030)
  for(const ExternalBinding &ext : m.externals)
  {
    if(ext.storageClass == SC_Output)
      continue;
    uint8_t *val = globalPtr(st, m, ext.var);
    if(ext.decoration.dec == Dec_BuiltIn)
    {
      uint32_t b = ext.decoration.param;
      if(b == BI_VertexIndex || b == BI_VertexId)
        memcpy(val, &vertexIndex, 4);
      else if(b == BI_InstanceIndex || b == BI_InstanceId)
        memset(val, 0, 4);
      else
      {
        fprintf(stderr, "vor: Unsupported builtin input\n");
        abort();
      }
    }
    else if(ext.decoration.dec == Dec_Location)
    {
      float a[4];
      env.vertex_attr(env.user, vertexIndex, ext.decoration.param, a);
      const Type &inner = m.types[globalPointee(m, ext.var)];
      if(inner.kind == T_VEC)
        memcpy(val, a, 4 * inner.count);    // bitcast for ints, truncating shuffle (:1971-1981)
      else
        memcpy(val, a, 4);
    }
    else
      bindResources(m, st, env, ext, false);
  }

  Interp in{m, st, env};
  in.exec(m.funcs.at(e->func), NULL, NULL);

  // outputs (:2039-2112)
  float *outpos = out;
  float *interp = out + 4;
  for(const ExternalBinding &ext : m.externals)
  {
    if(ext.storageClass != SC_Output)
      continue;
    uint8_t *ptr = globalPtr(st, m, ext.var);
    uint32_t tid = globalPointee(m, ext.var);
    if(ext.decoration.member != ~0U)
    {
      const Type &stt = m.types[tid];
      if(stt.kind != T_STRUCT || ext.decoration.member >= stt.members.size())
        continue;
      ptr += stt.offsets[ext.decoration.member];
      tid = stt.members[ext.decoration.member];
    }
    const Type &t = m.types[tid];
    if(ext.decoration.dec == Dec_Location)
    {
      float *dst = interp + 4 * ext.decoration.param;
      if(t.kind == T_VEC)
      {
        float v[4];
        memcpy(v, ptr, 4 * t.count);
        for(uint32_t i = t.count; i < 4; i++)
          v[i] = v[0];    // mask[i] = 0 (:2061-2064)
        memcpy(dst, v, 16);
      }
      else if(t.kind == T_FLOAT || t.kind == T_INT)
      {
        float s;
        memcpy(&s, ptr, 4);    // int: bitcast (:2072)
        dst[0] = dst[1] = dst[2] = dst[3] = s;
      }
      else if(t.kind == T_ARR || t.kind == T_MAT)
      {
        const Type &et = m.types[t.elem];
        for(uint32_t a = 0; a < t.count; a++)    // consecutive slots (:2079-2087)
          memcpy(dst + 4 * a, ptr + a * et.size, et.kind == T_VEC ? 4 * et.count : 4);
      }
    }
    else if(ext.decoration.dec == Dec_BuiltIn)
    {
      switch(ext.decoration.param)
      {
        case BI_Position: memcpy(outpos, ptr, 16); break;
        case BI_PointSize:
        case BI_ClipDistance:
        case BI_CullDistance: break;
        default: fprintf(stderr, "vor: Unsupported builtin output\n"); abort();
      }
    }
  }
}

bool run_fragment(const Entry *e, const ShaderEnv &env, float pixdepth, const float bary[4],
                  const float *tri, float out[4])
{
  (void)pixdepth;
  const Module &m = *e->mod;
  State &st = stateFor(&m);
  st.top = 0;
  st.killed = false;

  auto interps = [&](int vert, uint32_t loc) { return tri + vert * kVertexFloats + 4 + 4 * loc; };
  // CreateDot(loadedBary, (v0,v1,v2,0), 4) (:2202-2211)
  auto dot4 = [&](float a, float b, float c) {
    return ((bary[0] * a + bary[1] * b) + bary[2] * c) + bary[3] * 0.0f;
  };

  for(const ExternalBinding &ext : m.externals)
  {
    if(ext.storageClass == SC_Output)
      continue;
    uint8_t *val = globalPtr(st, m, ext.var);
    if(ext.decoration.dec == Dec_BuiltIn)
    {
      fprintf(stderr, "vor: Unsupported builtin input\n");    // :2155
      abort();
    }
    else if(ext.decoration.dec == Dec_Location)
    {
      uint32_t tid = globalPointee(m, ext.var);
      const Type &t = m.types[tid];
      uint32_t loc = ext.decoration.param;
      if(t.kind == T_ARR || t.kind == T_MAT || t.kind == T_VEC)
      {
        bool isArray = t.kind != T_VEC;
        uint32_t arraySize = isArray ? t.count : 1;
        const Type &vt = isArray ? m.types[t.elem] : t;
        for(uint32_t a = 0; a < arraySize; a++)
        {
          float v[4] = {0, 0, 0, 0};
          for(uint32_t i = 0; i < vt.count; i++)
            v[i] = dot4(interps(0, loc + a)[i], interps(1, loc + a)[i], interps(2, loc + a)[i]);
          memcpy(val + a * vt.size, v, 4 * vt.count);
        }
      }
      else if(t.kind == T_INT)
      {
        memcpy(val, interps(0, loc), 4);    // flat from vertex 0, bitcast (:2229-2247)
      }
      else
      {
        float s = dot4(interps(0, loc)[0], interps(1, loc)[0], interps(2, loc)[0]);
        memcpy(val, &s, 4);
      }
    }
    else
      bindResources(m, st, env, ext, true);
  }

  Interp in{m, st, env};
  in.exec(m.funcs.at(e->func), NULL, NULL);
  if(st.killed)
    return true;

  for(const ExternalBinding &ext : m.externals)
  {
    if(ext.storageClass != SC_Output)
      continue;
    uint8_t *ptr = globalPtr(st, m, ext.var);
    uint32_t tid = globalPointee(m, ext.var);
    if(ext.decoration.member != ~0U)
    {
      const Type &stt = m.types[tid];
      if(stt.kind != T_STRUCT || ext.decoration.member >= stt.members.size())
        continue;
      ptr += stt.offsets[ext.decoration.member];
    }
    if(ext.decoration.dec == Dec_Location)
    {
      // assert(param == 0); store val -> out (:2352-2354)
      memcpy(out, ptr, 16);
    }
    else if(ext.decoration.dec == Dec_BuiltIn)
    {
      fprintf(stderr, "vor: Unsupported builtin output\n");
      abort();
    }
  }
  return false;
}
}    // namespace vor
