// ptx_inline.cpp — textual inlining of PTX device functions at their call sites.
//
// ptxas does not inline `.func` calls: a shader function spliced into a kernel module stays a CALL/RET pair
// with register shuffling around it, its loads cannot be scheduled across the call, and every value that is
// live across it is pinned. The reference gets a fully inlined pixel loop from LLVM (AlwaysInline,
// spirv_compile.cpp:1143); here the same is done on the PTX text before ptxas sees it:
//
//   { // callseq                               { // inlined NAME #k
//   .param .b64 param0;                          .reg .b32 %i3_r<95>; ...          (the callee's registers, renamed)
//   st.param.b64 [param0], %rd151;       ->      mov.b64 %i3_rd0, %rd151;         (was ld.param [NAME_param_0])
//   ...                                          ... body, labels renamed ...
//   call.uni (retval0), NAME, (param0, ...);     mov.b32 %f196, %i3_r10; ...      (was st.param [func_retval0])
//   ld.param.v4.f32 {%f196, ...}, [retval0];     $i3_END:
//   }                                          }
//
// Handles the call sequences nvcc emits and the ones spirv_ptx.cpp emits, callees that nvcc compiled
// (vb200_sample_tex, vb200_fetch_attr) and callees from spirv_ptx.cpp (vb200_vs, vb200_fs). A callee that
// uses a stack frame (%SP / __local_depot) or a call site whose shape is not recognised is left as a call.
#include "ptx_inline.h"
#include <ctype.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <vector>

namespace vb200
{
namespace
{
struct Func
{
  std::string name;
  std::vector<std::string> params;    // parameter names in order
  std::string retName;                // "" when the function returns nothing
  std::string body;                   // everything between the braces (declarations included)
};

bool isIdent(char c)
{
  return isalnum((unsigned char)c) || c == '_' || c == '$' || c == '%';
}

// whole-word search
size_t findWord(const std::string &t, const std::string &w, size_t from)
{
  for(size_t at = t.find(w, from); at != std::string::npos; at = t.find(w, at + 1))
  {
    const bool l = at == 0 || !isIdent(t[at - 1]);
    const bool r = at + w.size() >= t.size() || !isIdent(t[at + w.size()]);
    if(l && r)
      return at;
  }
  return std::string::npos;
}

std::string trim(const std::string &s)
{
  size_t a = 0, b = s.size();
  while(a < b && isspace((unsigned char)s[a]))
    a++;
  while(b > a && isspace((unsigned char)s[b - 1]))
    b--;
  return s.substr(a, b - a);
}

// names declared by `.param <type...> NAME[dims]` items of a comma-separated list
std::vector<std::string> paramNames(const std::string &list)
{
  std::vector<std::string> out;
  size_t at = 0;
  while(at < list.size())
  {
    size_t end = list.find(',', at);
    if(end == std::string::npos)
      end = list.size();
    std::string item = trim(list.substr(at, end - at));
    at = end + 1;
    if(item.empty())
      continue;
    const size_t br = item.find('[');
    if(br != std::string::npos)
      item.erase(br);
    const size_t sp = item.find_last_of(" \t");
    out.push_back(sp == std::string::npos ? item : item.substr(sp + 1));
  }
  return out;
}

// definition `.func (ret) NAME (params) { body }` inside `text`
bool parseFunc(const std::string &text, const std::string &name, Func &f, size_t *defBegin = nullptr, size_t *defEnd = nullptr)
{
  for(size_t at = findWord(text, name, 0); at != std::string::npos; at = findWord(text, name, at + 1))
  {
    // the directive this occurrence belongs to: the nearest ".func" before it, on the way only a return list
    const size_t fn = text.rfind(".func", at);
    if(fn == std::string::npos)
      continue;
    const std::string between = text.substr(fn + 5, at - fn - 5);
    if(between.find(';') != std::string::npos || between.find('{') != std::string::npos ||
       between.find('}') != std::string::npos)
      continue;
    size_t lineStart = text.rfind('\n', fn);
    lineStart = lineStart == std::string::npos ? 0 : lineStart + 1;
    if(text.compare(lineStart, fn - lineStart, ".extern ") == 0)
      continue;
    size_t p = at + name.size();
    while(p < text.size() && isspace((unsigned char)text[p]))
      p++;
    if(p >= text.size() || text[p] != '(')
      continue;
    const size_t pe = text.find(')', p);
    if(pe == std::string::npos)
      continue;
    size_t b = pe + 1;
    while(b < text.size() && isspace((unsigned char)text[b]))
      b++;
    if(b >= text.size() || text[b] != '{')
      continue;    // a prototype
    // matching brace
    int depth = 0;
    size_t e = b;
    for(; e < text.size(); e++)
    {
      if(text[e] == '{')
        depth++;
      else if(text[e] == '}' && --depth == 0)
        break;
    }
    if(e >= text.size())
      return false;
    f.name = name;
    f.params = paramNames(text.substr(p + 1, pe - p - 1));
    f.retName.clear();
    const size_t ro = between.find('('), rc = between.rfind(')');
    if(ro != std::string::npos && rc != std::string::npos && rc > ro)
    {
      std::vector<std::string> r = paramNames(between.substr(ro + 1, rc - ro - 1));
      if(r.size() == 1)
        f.retName = r[0];
      else if(!r.empty())
        return false;
    }
    f.body = text.substr(b + 1, e - b - 1);
    if(defBegin)
      *defBegin = lineStart;
    if(defEnd)
      *defEnd = e + 1;
    return true;
  }
  return false;
}

struct Line
{
  std::string text;    // one statement (may hold several `;`-separated declarations), no trailing newline
};

std::vector<std::string> splitLines(const std::string &s)
{
  std::vector<std::string> out;
  size_t at = 0;
  while(at <= s.size())
  {
    size_t e = s.find('\n', at);
    if(e == std::string::npos)
      e = s.size();
    out.push_back(s.substr(at, e - at));
    at = e + 1;
  }
  return out;
}

// "[NAME+16]" / "[NAME]" -> NAME, offset
bool parseAddr(const std::string &s, std::string &name, unsigned &off)
{
  const size_t a = s.find('['), b = s.find(']');
  if(a == std::string::npos || b == std::string::npos || b < a)
    return false;
  std::string in = trim(s.substr(a + 1, b - a - 1));
  off = 0;
  const size_t plus = in.find('+');
  if(plus != std::string::npos)
  {
    off = (unsigned)strtoul(in.c_str() + plus + 1, nullptr, 0);
    in = trim(in.substr(0, plus));
  }
  name = in;
  return true;
}

// "{a, b, c}" or "a" -> operands
std::vector<std::string> operands(const std::string &s)
{
  std::string t = trim(s);
  if(!t.empty() && t[0] == '{')
  {
    const size_t e = t.find('}');
    t = t.substr(1, e == std::string::npos ? std::string::npos : e - 1);
  }
  std::vector<std::string> out;
  size_t at = 0;
  while(at <= t.size())
  {
    size_t e = t.find(',', at);
    if(e == std::string::npos)
      e = t.size();
    std::string o = trim(t.substr(at, e - at));
    if(!o.empty())
      out.push_back(o);
    at = e + 1;
  }
  return out;
}

// element size in bytes of the type suffix of a ld/st.param opcode ("ld.param.v4.f32" -> 4)
unsigned elemBytes(const std::string &opcode)
{
  const size_t dot = opcode.rfind('.');
  const std::string t = dot == std::string::npos ? "" : opcode.substr(dot + 1);
  if(t.size() >= 2 && isdigit((unsigned char)t[1]))
    return (unsigned)atoi(t.c_str() + 1) / 8u;
  return 0;
}

// the callee's registers and labels get a prefix of their own
std::string renameBody(const std::string &body, const std::string &tag)
{
  std::string out;
  out.reserve(body.size() + body.size() / 4);
  for(size_t i = 0; i < body.size();)
  {
    const char c = body[i];
    if(c == '%' && i + 1 < body.size() && isalpha((unsigned char)body[i + 1]))
    {
      size_t j = i + 1;
      while(j < body.size() && isalpha((unsigned char)body[j]))
        j++;
      // %name<digits> (a use) or %name< (a declaration): a general-purpose register; anything else
      // (%tid.x, %laneid, %SP ...) is a special register and stays
      if(j < body.size() && (isdigit((unsigned char)body[j]) || body[j] == '<'))
      {
        out += "%" + tag + "_";
        out.append(body, i + 1, j - i - 1);
        i = j;
        continue;
      }
    }
    if(c == '$' && i + 1 < body.size() && body[i + 1] == 'L' && (i == 0 || !isIdent(body[i - 1])))
    {
      // labels: $L... ; the debug-info strings nvcc references from .loc ($L__info_string<n>) are globals
      if(body.compare(i, 15, "$L__info_string") == 0)
      {
        out += c;
        i++;
        continue;
      }
      out += "$" + tag + "_";
      i++;
      continue;
    }
    out += c;
    i++;
  }
  return out;
}

struct CallSite
{
  size_t begin, end;                           // the `{ ... }` block, whole lines
  std::map<std::string, std::string> args;     // staging parameter -> value stored into it
  std::vector<std::string> argOrder;           // staging parameters in call order
  std::string retVar;
  std::map<unsigned, std::string> outs;        // byte offset inside the return value -> destination register
};

bool parseCallSite(const std::string &text, size_t callAt, const std::string &name, CallSite &cs)
{
  // "call.uni (rv), NAME, (a, b, c);" possibly spread over lines
  const size_t stmtEnd = text.find(';', callAt);
  if(stmtEnd == std::string::npos)
    return false;
  const std::string stmt = text.substr(callAt, stmtEnd - callAt);
  const size_t nm = findWord(stmt, name, 0);
  if(nm == std::string::npos)
    return false;
  const size_t ro = stmt.find('('), rc = stmt.find(')');
  cs.retVar.clear();
  size_t argOpen;
  if(ro != std::string::npos && ro < nm)
  {
    if(rc == std::string::npos || rc > nm)
      return false;
    cs.retVar = trim(stmt.substr(ro + 1, rc - ro - 1));
    argOpen = stmt.find('(', nm);
  }
  else
    argOpen = stmt.find('(', nm);
  if(argOpen == std::string::npos)
    cs.argOrder.clear();
  else
  {
    const size_t argClose = stmt.find(')', argOpen);
    if(argClose == std::string::npos)
      return false;
    cs.argOrder = operands(stmt.substr(argOpen + 1, argClose - argOpen - 1));
  }
  // the block around the call: the nearest line before it that is just "{" (a comment may follow), the first
  // line after it that starts with "}"
  size_t open = std::string::npos;
  size_t ls = text.rfind('\n', callAt);    // newline in front of the call's own line
  while(ls != std::string::npos && ls > 0)
  {
    const size_t prev = text.rfind('\n', ls - 1);
    const size_t start = prev == std::string::npos ? 0 : prev + 1;
    const std::string l = trim(text.substr(start, ls - start));
    if(!l.empty() && l[0] == '{' && (l.size() == 1 || trim(l.substr(1)).compare(0, 2, "//") == 0))
    {
      open = start;
      break;
    }
    if(!l.empty() && (l[0] == '}' || l.compare(0, 4, "call") == 0))
      return false;    // walked out of the sequence
    if(prev == std::string::npos)
      break;
    ls = prev;
  }
  if(open == std::string::npos)
    return false;
  size_t close = std::string::npos;
  for(size_t ls = text.find('\n', stmtEnd); ls != std::string::npos; ls = text.find('\n', ls + 1))
  {
    const size_t le = text.find('\n', ls + 1);
    const std::string l = trim(text.substr(ls + 1, (le == std::string::npos ? text.size() : le) - ls - 1));
    if(!l.empty() && l[0] == '}')
    {
      close = le == std::string::npos ? text.size() : le + 1;
      break;
    }
    if(l.find("call") == 0 || (!l.empty() && l[0] == '{'))
      return false;
  }
  if(close == std::string::npos)
    return false;
  cs.begin = open;
  cs.end = close;
  // statements of the block
  cs.args.clear();
  cs.outs.clear();
  const std::string head = text.substr(open, callAt - open), tail = text.substr(stmtEnd + 1, close - stmtEnd - 1);
  for(const std::string &raw : splitLines(head))
  {
    // a line may hold several statements
    size_t at = 0;
    while(at < raw.size())
    {
      size_t e = raw.find(';', at);
      if(e == std::string::npos)
        e = raw.size();
      const std::string st = trim(raw.substr(at, e - at));
      at = e + 1;
      if(st.compare(0, 9, "st.param.") != 0)
        continue;
      const size_t sp = st.find_first_of(" \t");
      const size_t comma = st.find(',', st.find(']'));
      if(sp == std::string::npos || comma == std::string::npos)
        return false;
      std::string pn;
      unsigned off;
      if(!parseAddr(st, pn, off) || off != 0)
        return false;
      std::vector<std::string> v = operands(st.substr(comma + 1));
      if(v.size() != 1)
        return false;    // (vector arguments are not used by any callee here)
      cs.args[pn] = v[0];
    }
  }
  for(const std::string &raw : splitLines(tail))
  {
    const std::string st = trim(raw);
    if(st.compare(0, 9, "ld.param.") != 0)
      continue;
    const size_t sp = st.find_first_of(" \t");
    const size_t br = st.find('[');
    if(sp == std::string::npos || br == std::string::npos)
      return false;
    const std::string opcode = st.substr(0, sp);
    const size_t comma = st.rfind(',', br);
    if(comma == std::string::npos)
      return false;
    std::string vn;
    unsigned off;
    if(!parseAddr(st, vn, off) || vn != cs.retVar)
      return false;
    const unsigned eb = elemBytes(opcode);
    if(!eb)
      return false;
    std::vector<std::string> dst = operands(st.substr(sp, comma - sp));
    for(size_t i = 0; i < dst.size(); i++)
      cs.outs[off + (unsigned)i * eb] = dst[i];
  }
  for(const std::string &a : cs.argOrder)
    if(!cs.args.count(a))
      return false;
  return true;
}

// register family (name without its number) -> width in bytes, from the `.reg .type %name<n>;` declarations
std::map<std::string, unsigned> regWidths(const std::string &body)
{
  std::map<std::string, unsigned> w;
  for(size_t at = body.find(".reg"); at != std::string::npos; at = body.find(".reg", at + 4))
  {
    const size_t semi = body.find(';', at);
    if(semi == std::string::npos)
      break;
    const std::string d = body.substr(at + 4, semi - at - 4);
    const size_t dot = d.find('.'), pc = d.find('%');
    if(dot == std::string::npos || pc == std::string::npos)
      continue;
    size_t te = dot + 1;
    while(te < d.size() && isalnum((unsigned char)d[te]))
      te++;
    const std::string type = d.substr(dot + 1, te - dot - 1);
    unsigned bytes = 0;
    if(type == "pred")
      bytes = 1;
    else if(type.size() >= 2 && isdigit((unsigned char)type[1]))
      bytes = (unsigned)atoi(type.c_str() + 1) / 8u;
    size_t ne = pc + 1;
    while(ne < d.size() && (isalnum((unsigned char)d[ne]) || d[ne] == '_') )
      ne++;
    std::string fam = d.substr(pc, ne - pc);
    if(d.find('<', pc) != std::string::npos)    // %name<n>: the family is the name as written
      w[fam] = bytes;
    else
      w[fam] = bytes;                           // a single named register
  }
  return w;
}

unsigned regBytes(const std::map<std::string, unsigned> &w, const std::string &reg)
{
  auto it = w.find(reg);
  if(it != w.end())
    return it->second;
  size_t e = reg.size();
  while(e > 0 && isdigit((unsigned char)reg[e - 1]))
    e--;
  it = w.find(reg.substr(0, e));
  return it == w.end() ? 0u : it->second;
}

// the callee's body rewritten for one call site, or "" when something in it is not understood
std::string instantiate(const Func &f, const CallSite &cs, const std::string &tag)
{
  if(f.body.find("%SP") != std::string::npos || f.body.find("__local_depot") != std::string::npos)
    return "";
  if(cs.argOrder.size() != f.params.size())
    return "";
  std::map<std::string, std::string> argOf;
  for(size_t i = 0; i < f.params.size(); i++)
    argOf[f.params[i]] = cs.args.at(cs.argOrder[i]);
  const std::string endLabel = "$" + tag + "_END";
  std::string out = "\t{ // inlined " + f.name + " " + tag + "\n";
  const std::string renamed = renameBody(f.body, tag);
  const std::map<std::string, unsigned> widths = regWidths(renamed);
  std::map<unsigned, bool> written;
  for(const std::string &raw : splitLines(renamed))
  {
    const std::string st = trim(raw);
    if(st.compare(0, 9, "ld.param.") == 0)
    {
      const size_t sp = st.find_first_of(" \t"), br = st.find('['), comma = st.rfind(',', br);
      std::string pn;
      unsigned off;
      if(sp == std::string::npos || br == std::string::npos || comma == std::string::npos || !parseAddr(st, pn, off))
        return "";
      if(!argOf.count(pn))
      {
        out += raw + "\n";    // the return value of a nested call
        continue;
      }
      if(off != 0)
        return "";
      const std::string opcode = st.substr(0, sp);
      const unsigned eb = elemBytes(opcode);
      std::vector<std::string> dst = operands(st.substr(sp, comma - sp));
      if(dst.size() != 1 || (eb != 4 && eb != 8))
        return "";
      // nvcc loads 32-bit parameters straight into 64-bit registers where the value is only used as an offset
      // (the load extends): the move has to extend as well
      const unsigned dw = regBytes(widths, dst[0]);
      const std::string &arg = argOf[pn];
      if(dw == 8 && eb == 4)
      {
        if(!arg.empty() && arg[0] == '%')
          out += std::string("\tcvt.") + (opcode.find(".s32") != std::string::npos ? "s64.s32 " : "u64.u32 ") + dst[0] + ", " + arg + ";\n";
        else
          out += "\tmov.b64 " + dst[0] + ", " + arg + ";\n";
      }
      else if(dw && dw != eb)
        return "";
      else
        out += "\tmov.b" + std::to_string(eb * 8) + " " + dst[0] + ", " + arg + ";\n";
      continue;
    }
    if(st.compare(0, 9, "st.param.") == 0)
    {
      const size_t sp = st.find_first_of(" \t"), br = st.find('['), close = st.find(']'), semi = st.rfind(';');
      std::string pn;
      unsigned off;
      if(sp == std::string::npos || br == std::string::npos || close == std::string::npos || !parseAddr(st, pn, off))
        return "";
      if(pn != f.retName)
      {
        out += raw + "\n";    // argument staging of a nested call
        continue;
      }
      const unsigned eb = elemBytes(st.substr(0, sp));
      const size_t comma = st.find(',', close);
      if(comma == std::string::npos || (eb != 4 && eb != 8))
        return "";
      std::vector<std::string> src =
          operands(st.substr(comma + 1, (semi == std::string::npos ? st.size() : semi) - comma - 1));
      for(size_t i = 0; i < src.size(); i++)
      {
        auto it = cs.outs.find(off + (unsigned)i * eb);
        if(it != cs.outs.end())
        {
          out += "\tmov.b" + std::to_string(eb * 8) + " " + it->second + ", " + src[i] + ";\n";
          written[it->first] = true;
        }
      }
      continue;
    }
    if(st == "ret;")
    {
      out += "\tbra " + endLabel + ";\n";
      continue;
    }
    out += raw + "\n";
  }
  // words of the return value the callee never stores (padding) read as zero
  for(auto &kv : cs.outs)
    if(!written.count(kv.first))
      out += "\tmov.b32 " + kv.second + ", 0;\n";
  out += endLabel + ":\n\t}\n";
  return out;
}
}    // namespace

int ptx_inline_calls(std::string &text, const std::string &defs, const std::string &name, int *serial)
{
  Func f;
  if(!parseFunc(defs, name, f))
    return 0;
  int done = 0;
  size_t from = 0;
  for(;;)
  {
    // next "call" statement that names the function
    size_t callAt = std::string::npos;
    for(size_t at = text.find("call", from); at != std::string::npos; at = text.find("call", at + 4))
    {
      if(at > 0 && isIdent(text[at - 1]))
        continue;
      const size_t semi = text.find(';', at);
      if(semi == std::string::npos)
        break;
      if(findWord(text.substr(at, semi - at), name, 0) != std::string::npos)
      {
        callAt = at;
        break;
      }
    }
    if(callAt == std::string::npos)
      break;
    CallSite cs;
    std::string inst;
    if(parseCallSite(text, callAt, name, cs))
      inst = instantiate(f, cs, "i" + std::to_string((*serial)++));
    if(inst.empty())
    {
      from = callAt + 4;    // stays a call
      continue;
    }
    text.replace(cs.begin, cs.end - cs.begin, inst);
    from = cs.begin;    // the inlined body may contain calls to the same function's helpers, never to itself
    done++;
  }
  return done;
}

bool ptx_has_call(const std::string &text, const std::string &name)
{
  for(size_t at = text.find("call", 0); at != std::string::npos; at = text.find("call", at + 4))
  {
    if(at > 0 && isIdent(text[at - 1]))
      continue;
    const size_t semi = text.find(';', at);
    if(semi == std::string::npos)
      break;
    if(findWord(text.substr(at, semi - at), name, 0) != std::string::npos)
      return true;
  }
  return false;
}
}    // namespace vb200
