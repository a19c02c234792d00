/*
 * dsl.hpp — front end of the plan compiler: the reference's `graph!` DSL, as text, to a node plan.
 *
 * The reference compiles a `graph! { ... }` block at Rust compile time
 * (oscen-graph-compiler/src/lib.rs:22-34): parse (parse.rs) -> lower to IR with type checks
 * (ir/lower.rs:351-656) -> dead-node removal (ir/passes/dead_nodes.rs) -> Kahn sort
 * (ir/lower.rs:1015-1085) -> Rust code that calls every node's process() per frame.  Here the same
 * text is compiled at run time into a plan::Graph (plan.hpp): one node-kernel launch per node type
 * per ready wave, streams in HBM, replayed as a CUDA graph (kernel_nodes.cuh).
 *
 * Grammar accepted (parse.rs:195-1110): `name: X;`, `input`/`output` declarations in both
 * spellings (`input cutoff: value = 3000.0 [20.0..20000.0, log, ramp: 1323];` and
 * `input value cutoff = 3000.0;`), `node x = T::ctor(args);` / `nodes { ... }`,
 * `connection a -> b;` / `connections { ... }`, compound sources with + - * / and parentheses,
 * f32 methods tanh/abs/clamp/min/max, optional `[policy]` prefixes, an optional surrounding
 * `graph! { }`.  Semantics kept: kind compatibility (lower.rs:1157-1165: Stream->Stream,
 * Value->Value, Event->Event, Value->Stream), auto-summing stream fan-in in statement order
 * (codegen/emit_node.rs:153-174: connect the first source, accumulate the rest), last write wins
 * on value endpoints, constructor arguments stay on unconnected endpoints, nodes that do not
 * reach an output are dropped.
 *
 * Multirate: `node = T::new() * N` (N in {2,4,8}, one factor per graph) runs N times per frame at
 * sample rate sr * N; every stream edge that crosses the rate boundary gets its own resampler node
 * (Up / Down) of the edge's `[policy]` (default [sinc]; ir/lower.rs:741-906), value edges latch.
 * Multirate graphs run as generated kernels only (codegen.hpp), not as node plans.
 *
 * Feedback: `src -> [N] -> dst` (a hidden Delay::new(N, 0.0)) and `src -> [name] -> dst` (a declared Delay)
 * close cycles; the leg out of the delay does not order the walk (ir/lower.rs:560-656, 1015-1085) and reads
 * what the delay output when it last ran.  Cycles need the per-sample walk: generated kernels only.
 * Polyphony: a text may hold several `graph! { }` blocks; an outer graph of the reference's polyphonic
 * shape (voices = [Voice::new(); N] + the MIDI chain) lowers to the voice graph's plan (lower_polyphonic).
 *
 * Assets: `external NAME: AudioAsset;` declares a slot the host loads audio into at run time
 * (oscen_b200_load_asset); `NAME -> player.buf` binds it to a SamplePlayer.
 *
 * Node arrays of library types (`[T::new(); N]`, `nodes[k].ep`) are unrolled into scalar nodes with the reference's
 * Broadcast / Parallel / FanIn edge shapes (ArrayExpander).
 *
 * Reported as errors (never silently ignored): `/ N` rates, mixed factors, arrays of graphs outside the
 * polyphonic shape, externals other than AudioAsset, typed endpoints, free-function calls in
 * sources (they name user Rust code).
 *
 * Host-only C++, no CUDA types: unit-tested on the CPU through oscen_b200_graph_compile().
 */
#pragma once
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "plan.hpp"

namespace oscen {
namespace dsl {

/* ------------------------------------------------------------------------------ lexer -- */
struct Tok {
    enum Kind { Ident, Num, Str, Punct, End } kind = End;
    std::string s;
    float num = 0.0f;
    bool is_int = false;
    int line = 1, col = 1;
};

struct Error {
    std::string msg;
    int line = 0, col = 0;
    bool set = false;
    void at(const Tok &t, const std::string &m) { if (!set) { set = true; msg = m; line = t.line; col = t.col; } }
    std::string text() const { return std::to_string(line) + ":" + std::to_string(col) + ": " + msg; }
};

inline bool lex(const char *src, std::vector<Tok> &out, Error &err) {
    int line = 1, col = 1;
    const char *p = src;
    auto adv = [&](int n) { for (int i = 0; i < n; ++i) { if (*p == '\n') { ++line; col = 1; } else ++col; ++p; } };
    while (*p) {
        if (isspace((unsigned char)*p)) { adv(1); continue; }
        if (p[0] == '/' && p[1] == '/') { while (*p && *p != '\n') adv(1); continue; }
        if (p[0] == '/' && p[1] == '*') {
            Tok t; t.line = line; t.col = col;
            adv(2);
            while (*p && !(p[0] == '*' && p[1] == '/')) adv(1);
            if (!*p) { err.at(t, "unterminated block comment"); return false; }
            adv(2); continue;
        }
        Tok t; t.line = line; t.col = col;
        if (isalpha((unsigned char)*p) || *p == '_') {
            const char *q = p;
            while (isalnum((unsigned char)*q) || *q == '_') ++q;
            /* names travel into fixed 64-byte fields (oscen_b200_input_info.name) and into the generated kernel's text */
            if (q - p > 63) { err.at(t, "identifier longer than 63 characters"); return false; }
            t.kind = Tok::Ident; t.s.assign(p, q); adv((int)(q - p));
        } else if (isdigit((unsigned char)*p)) {
            /* Rust numeric literal: digits with `_`, optional fraction, exponent, type suffix.
               `20.0..20000.0` must not swallow the range dots. */
            const char *q = p;
            std::string txt;
            bool is_int = true;
            auto digits = [&]() { while (isdigit((unsigned char)*q) || *q == '_') { if (*q != '_') txt += *q; ++q; } };
            digits();
            if (q[0] == '.' && q[1] != '.' && !(isalpha((unsigned char)q[1]) || q[1] == '_')) { is_int = false; txt += '.'; ++q; digits(); }
            if ((*q == 'e' || *q == 'E') && (isdigit((unsigned char)q[1]) || ((q[1] == '+' || q[1] == '-') && isdigit((unsigned char)q[2])))) {
                is_int = false; txt += *q++; if (*q == '+' || *q == '-') txt += *q++; digits();
            }
            if (!strncmp(q, "f32", 3) || !strncmp(q, "f64", 3)) { is_int = false; q += 3; }
            else if (!strncmp(q, "usize", 5)) q += 5;
            else if (!strncmp(q, "u32", 3) || !strncmp(q, "i32", 3) || !strncmp(q, "u64", 3) || !strncmp(q, "i64", 3)) q += 3;
            t.kind = Tok::Num; t.s = txt; t.is_int = is_int;
            t.num = strtof(txt.c_str(), nullptr);        /* decimal -> nearest f32, as rustc rounds an f32 literal */
            adv((int)(q - p));
        } else if (*p == '"') {
            const char *q = p + 1;
            while (*q && *q != '"') { if (*q == '\\' && q[1]) ++q; ++q; }
            if (!*q) { err.at(t, "unterminated string literal"); return false; }
            t.kind = Tok::Str; t.s.assign(p + 1, q); adv((int)(q - p) + 1);
        } else {
            t.kind = Tok::Punct;
            if ((p[0] == '-' && p[1] == '>') || (p[0] == ':' && p[1] == ':') || (p[0] == '.' && p[1] == '.')) { t.s.assign(p, 2); adv(2); }
            else if (strchr("{}[]();:,=-+*/.<>!", *p)) { t.s.assign(p, 1); adv(1); }
            else {
                char shown[8];
                if ((unsigned char)*p >= 0x20 && (unsigned char)*p < 0x7f) snprintf(shown, sizeof shown, "%c", *p);
                else snprintf(shown, sizeof shown, "\\x%02x", (unsigned)(unsigned char)*p);      /* keep the diagnostic valid UTF-8 */
                err.at(t, std::string("unexpected character `") + shown + "`"); return false;
            }
        }
        out.push_back(t);
    }
    Tok e; e.kind = Tok::End; e.line = line; e.col = col; out.push_back(e);
    return true;
}

/* ---------------------------------------------------------- oversample_variants! -- */
/* `oversample_variants! { base_name: X; factors: [1, 4]; body: { ... {FACTOR} ... } }` stamps out one graph per
   factor, named X_1x, X_4x, ..., with every `{FACTOR}` replaced by the factor
   (oscen-macros/src/oversample_variants_macro.rs:94-139).  Done on the token stream, before parsing. */
inline bool expand_oversample_variants(std::vector<Tok> &toks, Error &err) {
    auto is = [&](size_t i, Tok::Kind k, const char *s) { return i < toks.size() && toks[i].kind == k && toks[i].s == s; };
    std::vector<Tok> out;
    size_t i = 0;
    while (i < toks.size()) {
        if (!(is(i, Tok::Ident, "oversample_variants") && is(i + 1, Tok::Punct, "!") && is(i + 2, Tok::Punct, "{"))) { out.push_back(toks[i++]); continue; }
        const Tok at = toks[i];
        size_t j = i + 3;
        std::string base; std::vector<int> factors; size_t body_lo = 0, body_hi = 0;
        while (j < toks.size() && !is(j, Tok::Punct, "}")) {
            if (toks[j].kind != Tok::Ident || !is(j + 1, Tok::Punct, ":")) { err.at(toks[j], "oversample_variants!: expected `base_name:`, `factors:` or `body:`"); return false; }
            const std::string key = toks[j].s; j += 2;
            if (key == "base_name") {
                if (j >= toks.size() || toks[j].kind != Tok::Ident) { err.at(toks[j < toks.size() ? j : toks.size() - 1], "oversample_variants!: `base_name:` expects an identifier"); return false; }
                base = toks[j++].s;
            } else if (key == "factors") {
                if (!is(j, Tok::Punct, "[")) { err.at(toks[j], "oversample_variants!: `factors:` expects `[1, 2, ...]`"); return false; }
                for (++j; j < toks.size() && !is(j, Tok::Punct, "]"); ++j) {
                    if (is(j, Tok::Punct, ",")) continue;
                    if (toks[j].kind != Tok::Num || !toks[j].is_int) { err.at(toks[j], "oversample_variants!: factors are integer literals"); return false; }
                    factors.push_back((int)strtol(toks[j].s.c_str(), nullptr, 10));
                }
                ++j;
            } else if (key == "body") {
                if (!is(j, Tok::Punct, "{")) { err.at(toks[j], "oversample_variants!: `body:` expects `{ ... }`"); return false; }
                int depth = 0; body_lo = j + 1;
                for (; j < toks.size(); ++j) {
                    if (is(j, Tok::Punct, "{")) ++depth;
                    else if (is(j, Tok::Punct, "}") && --depth == 0) break;
                }
                if (j >= toks.size()) { err.at(at, "oversample_variants!: unterminated body"); return false; }
                body_hi = j++;
            } else { err.at(toks[j - 2], "oversample_variants!: unknown key `" + key + "`"); return false; }
            if (is(j, Tok::Punct, ";") || is(j, Tok::Punct, ",")) ++j;
        }
        if (j >= toks.size()) { err.at(at, "oversample_variants!: missing `}`"); return false; }
        if (base.empty() || factors.empty() || !body_hi) { err.at(at, "oversample_variants! needs base_name, factors and body"); return false; }
        for (int f : factors) {
            auto push = [&](Tok::Kind k, const std::string &s) { Tok t = at; t.kind = k; t.s = s; out.push_back(t); };
            push(Tok::Ident, "graph"); push(Tok::Punct, "!"); push(Tok::Punct, "{");
            push(Tok::Ident, "name"); push(Tok::Punct, ":"); push(Tok::Ident, base + "_" + std::to_string(f) + "x"); push(Tok::Punct, ";");
            for (size_t k = body_lo; k < body_hi; ++k) {
                if (is(k, Tok::Punct, "{") && is(k + 1, Tok::Ident, "FACTOR") && is(k + 2, Tok::Punct, "}")) {
                    Tok t = toks[k]; t.kind = Tok::Num; t.s = std::to_string(f); t.num = (float)f; t.is_int = true; out.push_back(t);
                    k += 2;
                } else out.push_back(toks[k]);
            }
            push(Tok::Punct, "}");
        }
        i = j + 1;
        if (is(i, Tok::Punct, ";")) ++i;
    }
    toks.swap(out);
    return true;
}

/* -------------------------------------------------------------------------------- AST -- */
enum class Kind { Stream, Value, Event, Asset };
inline const char *kind_name(Kind k) { return k == Kind::Stream ? "Stream" : k == Kind::Value ? "Value" : k == Kind::Event ? "Event" : "Asset"; }

struct Expr {
    enum K { Ident, Field, Method, Binary, Literal, Call } k = Ident;
    std::string name;                      /* Ident: name; Field: field; Method: method; Call: function path */
    std::unique_ptr<Expr> a, b;            /* Field/Method: a = receiver; Binary: a (op) b */
    char op = 0;
    float lit = 0.0f;
    std::vector<float> args;               /* Method: literal arguments */
    Tok at;
};
using ExprP = std::unique_ptr<Expr>;

struct InputDecl { Kind kind; std::string name; bool has_def = false; float def = 0.0f; plan::ValueInput spec; Tok at; };
struct OutputDecl { Kind kind; std::string name; Tok at; int lanes = 1;   /* `stream: Frame<N>` / `stream: [f32; N]` (parse.rs endpoint types) */ };
struct NodeDecl { std::string name, type, ctor; std::vector<float> args; bool args_ok = true; bool bare = false; int array = 0; int rate = 0; Tok at; };
struct Conn { std::string policy; ExprP src, dst; int via = 0; /* 0 none, 1 samples, 2 node */ float via_samples = 0.0f; std::string via_node; Tok at; };
struct ExternalDecl { std::string name, type; Tok at; };     /* `external sample: AudioAsset;` (parse.rs:447-456) */
struct Ast {
    std::string name;
    std::vector<ExternalDecl> externals;
    std::vector<InputDecl> inputs;
    std::vector<OutputDecl> outputs;
    std::vector<NodeDecl> nodes;
    std::vector<Conn> conns;
};

/* ----------------------------------------------------------------------------- parser -- */
class Parser {
public:
    Parser(const std::vector<Tok> &t, Error &e) : toks(t), err(e) {}

    /* one graph body, or several `graph! { ... }` invocations in a row (a voice graph followed by the
       polyphonic graph that instantiates it, like fm_voice.rs + lib.rs of the fm-synth example) */
    bool parse(std::vector<Ast> &graphs) {
        auto wrapper = [&]() { return is_ident("graph") && peek(1).kind == Tok::Punct && peek(1).s == "!" && peek(2).s == "{"; };
        if (!wrapper()) {
            graphs.emplace_back();
            while (!err.set && cur().kind != Tok::End) item(graphs.back());
            return !err.set;
        }
        while (!err.set && wrapper()) {
            pos += 3;
            graphs.emplace_back();
            while (!err.set && cur().kind != Tok::End && !is_punct("}")) item(graphs.back());
            if (!err.set) expect("}");
            accept(";");
        }
        if (!err.set && cur().kind != Tok::End) err.at(cur(), "unexpected tokens after the graph body");
        return !err.set;
    }

private:
    const std::vector<Tok> &toks;
    Error &err;
    size_t pos = 0;

    const Tok &cur() const { return toks[pos]; }
    const Tok &peek(size_t n) const { return toks[std::min(pos + n, toks.size() - 1)]; }
    bool is_punct(const char *s) const { return cur().kind == Tok::Punct && cur().s == s; }
    bool is_ident(const char *s) const { return cur().kind == Tok::Ident && cur().s == s; }
    bool accept(const char *s) { if (is_punct(s)) { ++pos; return true; } return false; }
    bool expect(const char *s) {
        if (accept(s)) return true;
        err.at(cur(), std::string("expected `") + s + "`" + (cur().kind == Tok::End ? " before the end of the graph" : ", found `" + cur().s + "`"));
        return false;
    }
    bool ident(std::string &out, const char *what) {
        if (cur().kind != Tok::Ident) { err.at(cur(), std::string("expected ") + what); return false; }
        out = cur().s; ++pos; return true;
    }
    bool endpoint_kind(const std::string &s, Kind &k) {
        if (s == "stream") k = Kind::Stream; else if (s == "value") k = Kind::Value;
        else if (s == "event") k = Kind::Event; else if (s == "asset") k = Kind::Asset; else return false;
        return true;
    }
    /* skip a balanced token run up to (not including) a top-level `;` — type annotations etc. */
    void skip_to_semicolon() {
        int depth = 0;
        while (cur().kind != Tok::End) {
            if (cur().kind == Tok::Punct) {
                if (cur().s == "{" || cur().s == "[" || cur().s == "(") ++depth;
                else if (cur().s == "}" || cur().s == "]" || cur().s == ")") --depth;
                else if (cur().s == ";" && depth <= 0) return;
            }
            ++pos;
        }
    }
    bool number(float &v) {
        bool neg = false;
        if (is_punct("-")) { neg = true; ++pos; }
        if (cur().kind != Tok::Num) { err.at(cur(), "expected a numeric literal"); return false; }
        v = neg ? -cur().num : cur().num; ++pos;
        return true;
    }

    void item(Ast &ast) {
        const Tok t = cur();
        if (t.kind != Tok::Ident) { err.at(t, "expected `input`, `output`, `node`, `nodes`, `connection` or `connections`"); return; }
        if (t.s == "name" && peek(1).kind == Tok::Punct && peek(1).s == ":") {
            pos += 2;
            if (!ast.inputs.empty() || !ast.outputs.empty() || !ast.nodes.empty() || !ast.conns.empty() || !ast.name.empty()) {
                err.at(t, "`name:` must be the first item of the graph"); return;
            }
            if (ident(ast.name, "the graph name")) expect(";");
        } else if (t.s == "nih_params") { ++pos; expect(";");
        } else if (t.s == "input") { ++pos; input_decl(ast);
        } else if (t.s == "output") { ++pos; output_decl(ast);
        } else if (t.s == "external") {
            ++pos;
            ExternalDecl d; d.at = t;
            if (!ident(d.name, "the name of the external slot") || !expect(":") || !ident(d.type, "the slot's type (AudioAsset)")) return;
            if (expect(";")) ast.externals.push_back(d);
        } else if (t.s == "node" || t.s == "nodes") {
            ++pos;
            if (accept("{")) { while (!err.set && !is_punct("}") && cur().kind != Tok::End) { if (accept(";")) continue; node_decl(ast); } expect("}"); }
            else node_decl(ast);
        } else if (t.s == "connection" || t.s == "connections") {
            ++pos;
            if (accept("{")) { while (!err.set && !is_punct("}") && cur().kind != Tok::End) { if (accept(";")) continue; conn_stmt(ast); } expect("}"); }
            else conn_stmt(ast);
        } else err.at(t, "unknown graph item `" + t.s + "`");
    }

    /* parse.rs:386-445 */
    void input_decl(Ast &ast) {
        InputDecl d; d.at = cur();
        std::string first;
        if (!ident(first, "an input name or endpoint kind")) return;
        if (accept(":")) {
            std::string k;
            if (!ident(k, "an endpoint kind (stream, value, event)")) return;
            if (!endpoint_kind(k, d.kind)) { err.at(d.at, "unknown endpoint kind `" + k + "`"); return; }
            d.name = first;
        } else {
            if (!endpoint_kind(first, d.kind)) { err.at(d.at, "unknown endpoint kind `" + first + "`"); return; }
            if (!ident(d.name, "an input name")) return;
        }
        if (accept(":")) {                                   /* `: Type` annotation (array / frame payloads) */
            err.at(cur(), "typed endpoints (`: [f32; N]`, `: Frame<N>`) are not supported by the node plan: streams are mono f32");
            return;
        }
        if (accept("=")) {
            if (!number(d.def)) return;
            d.has_def = true;
            if (is_punct("[") || is_punct("{")) param_spec(d.spec);
        }
        d.spec.name = d.name; d.spec.def = d.def;
        if (expect(";")) ast.inputs.push_back(std::move(d));
    }
    /* `[min..max, log, ramp: N]` (ast.rs:175-187) or the brace form `{ range: a..b, unit: " Hz", ... }` (parse.rs:299-384);
       range / curve / ramp are kept, the NIH-plug display fields are skipped */
    void param_spec(plan::ValueInput &s) {
        const bool brace = is_punct("{");
        ++pos;
        const char *close = brace ? "}" : "]";
        while (!err.set && !is_punct(close) && cur().kind != Tok::End) {
            if (cur().kind == Tok::Num || is_punct("-")) {
                float lo, hi;
                if (!number(lo) || !expect("..") || !number(hi)) return;
                s.has_range = true; s.lo = lo; s.hi = hi;
            } else if (cur().kind == Tok::Ident) {
                const std::string key = cur().s; ++pos;
                if (key == "log") s.log = true;
                else if (key == "linear") s.log = false;
                else if (key == "range" && accept(":")) {
                    float lo, hi;
                    if (!number(lo) || !expect("..") || !number(hi)) return;
                    s.has_range = true; s.lo = lo; s.hi = hi;
                } else if (key == "ramp" && (accept(":") || accept("="))) {
                    if (cur().kind != Tok::Num || !cur().is_int) { err.at(cur(), "`ramp:` expects a frame count"); return; }
                    s.ramp = (uint32_t)strtoul(cur().s.c_str(), nullptr, 10); ++pos;
                } else if (accept(":") || accept("=")) {     /* center / unit / smoother / step / name / group (`key: v` or `key = v`) */
                    if (cur().kind == Tok::Str) ++pos; else { float v; if (!number(v)) return; }
                } else { err.at(cur(), "unknown parameter-spec entry `" + key + "`"); return; }
            } else { err.at(cur(), "malformed parameter spec"); return; }
            if (!accept(",")) break;
        }
        expect(close);
    }
    /* parse.rs:640-672 */
    void output_decl(Ast &ast) {
        OutputDecl d; d.at = cur();
        std::string first;
        if (!ident(first, "an output name or endpoint kind")) return;
        if (accept(":")) {
            std::string k;
            if (!ident(k, "an endpoint kind (stream, value, event)")) return;
            if (!endpoint_kind(k, d.kind)) { err.at(d.at, "unknown endpoint kind `" + k + "`"); return; }
            d.name = first;
        } else {
            if (!endpoint_kind(first, d.kind)) { err.at(d.at, "unknown endpoint kind `" + first + "`"); return; }
            if (!ident(d.name, "an output name")) return;
        }
        if (accept(":")) {                                   /* endpoint type: `f32`, `Frame<N>`, `[f32; N]` */
            const Tok at = cur();
            auto int_lit = [&](int &v) {
                if (cur().kind != Tok::Num || !cur().is_int) { err.at(cur(), "expected an integer literal"); return false; }
                v = (int)strtol(cur().s.c_str(), nullptr, 10); ++pos; return true;
            };
            if (accept("[")) {
                std::string el;
                if (!ident(el, "an element type") || el != "f32" || !expect(";") || !int_lit(d.lanes) || !expect("]")) { if (!err.set) err.at(at, "array-typed endpoints are `[f32; N]`"); return; }
            } else {
                std::string ty;
                if (!ident(ty, "an endpoint type (`f32`, `Frame<N>`, `[f32; N]`)")) return;
                if (ty == "Frame") { if (!expect("<") || !int_lit(d.lanes) || !expect(">")) return; }
                else if (ty != "f32") { err.at(at, "unknown endpoint type `" + ty + "` (`f32`, `Frame<N>`, `[f32; N]`)"); return; }
            }
            if (d.kind != Kind::Stream) { err.at(at, "only stream endpoints carry a frame type"); return; }
            if (d.lanes < 1 || d.lanes > 2) { err.at(at, "graph outputs are mono (`f32`, `Frame<1>`) or stereo (`Frame<2>`): the engine's instance and mix buses have 1 or 2 channels"); return; }
        }
        if (expect(";")) ast.outputs.push_back(std::move(d));
    }
    /* `<name> = <constructor> [* N | / N];` (parse.rs:679-710); constructor = `Type::method(args)`,
       `Type<G>::method(args)` or the array form `[ctor; N]` */
    void node_decl(Ast &ast) {
        NodeDecl d; d.at = cur();
        if (!ident(d.name, "a node name") || !expect("=")) return;
        const bool array = accept("[");
        std::string seg;
        if (!ident(seg, "a node constructor (`Type::new(...)`)")) return;
        std::vector<std::string> path{seg};
        for (;;) {
            if (is_punct("<")) {                             /* generic arguments: kept out of the type name */
                int depth = 0;
                do { if (is_punct("<")) ++depth; else if (is_punct(">")) --depth; ++pos; } while (depth > 0 && cur().kind != Tok::End);
            }
            if (!accept("::")) break;
            if (is_punct("<")) continue;                     /* turbofish */
            if (!ident(seg, "a path segment")) return;
            path.push_back(seg);
        }
        if (path.size() == 1 && !is_punct("(")) {             /* `voice = SimpleVoice;`: a graph used as a node */
            d.type = path[0]; d.ctor = "new"; d.bare = true;
        } else {
        if (path.size() < 2) { err.at(d.at, "node constructor must be `Type::method(...)`"); return; }
        d.ctor = path.back(); d.type = path[path.size() - 2];
        if (!expect("(")) return;
        while (!err.set && !is_punct(")") && cur().kind != Tok::End) {
            float v;
            if (cur().kind == Tok::Num || is_punct("-")) { if (!number(v)) return; d.args.push_back(v); }
            else { d.args_ok = false; int depth = 0;            /* a non-literal argument: reported when the node is lowered */
                   while (cur().kind != Tok::End && !(depth == 0 && (is_punct(",") || is_punct(")")))) {
                       if (is_punct("(") || is_punct("[")) ++depth; else if (is_punct(")") || is_punct("]")) --depth; ++pos; } }
            if (!accept(",")) break;
        }
        if (!expect(")")) return;
        }
        if (array) {
            if (!expect(";")) return;
            if (cur().kind != Tok::Num || !cur().is_int) { err.at(cur(), "array size must be an integer literal"); return; }
            d.array = (int)strtol(cur().s.c_str(), nullptr, 10); ++pos;
            if (!expect("]")) return;
        }
        if (is_punct("*") || is_punct("/")) {                /* parse.rs:460-488: power-of-two rate factor */
            const bool up = is_punct("*"); ++pos;
            if (cur().kind != Tok::Num || !cur().is_int) { err.at(cur(), "rate factor must be an integer literal in {1, 2, 4, 8}"); return; }
            const int n = (int)strtol(cur().s.c_str(), nullptr, 10);
            if (n != 1 && n != 2 && n != 4 && n != 8) { err.at(cur(), "rate factor must be in {1, 2, 4, 8}"); return; }
            ++pos;
            d.rate = n == 1 ? 0 : (up ? n : -n);
        }
        if (expect(";")) ast.nodes.push_back(std::move(d));
    }
    /* `[<policy>] <source> -> <dest>;` or `<source> -> [N | name] -> <dest>;` (parse.rs:931-980) */
    void conn_stmt(Ast &ast) {
        Conn c; c.at = cur();
        if (is_punct("[")) {
            ++pos;
            if (!ident(c.policy, "a connection policy (latch, linear, sinc, sinc_iir)")) return;
            if (c.policy != "latch" && c.policy != "linear" && c.policy != "sinc" && c.policy != "sinc_iir") { err.at(c.at, "unknown connection policy `" + c.policy + "`"); return; }
            if (!expect("]")) return;
        }
        expr_depth = 0; expr_ops = 0;
        c.src = additive();
        if (err.set || !expect("->")) return;
        if (is_punct("[")) {
            ++pos;
            if (cur().kind == Tok::Num) {
                if (!cur().is_int) { err.at(cur(), "the sample count of `-> [N] ->` must be an integer literal"); return; }
                c.via = 1; c.via_samples = cur().num;
            } else if (cur().kind == Tok::Ident) { c.via = 2; c.via_node = cur().s; }
            else { err.at(cur(), "expected a sample count or a node name inside `-> [ ] ->`"); return; }
            ++pos;
            if (!expect("]") || !expect("->")) return;
        }
        expr_depth = 0; expr_ops = 0;
        c.dst = additive();
        if (err.set) return;
        if (expect(";")) ast.conns.push_back(std::move(c));
    }
    /* parse.rs:989-1029: + - below * /, left associative; a `-` that starts `->` ends the expression */
    /* run-time text: the recursion (and the tree every later pass walks recursively) is bounded — a source nested
       deeper than kMaxExprDepth parentheses or longer than kMaxExprOps operators is a diagnostic, not a stack overflow */
    static constexpr int kMaxExprDepth = 64, kMaxExprOps = 1024;
    int expr_depth = 0, expr_ops = 0;
    bool count_op() {
        if (++expr_ops > kMaxExprOps) { err.at(cur(), "expression too long (more than 1024 operators in one connection)"); return false; }
        return true;
    }
    ExprP additive() {
        ExprP l = multiplicative();
        while (!err.set && (is_punct("+") || is_punct("-"))) {
            if (!count_op()) return l;
            ExprP e(new Expr); e->k = Expr::Binary; e->op = cur().s[0]; e->at = cur(); ++pos;
            e->a = std::move(l); e->b = multiplicative(); l = std::move(e);
        }
        return l;
    }
    ExprP multiplicative() {
        ExprP l = primary();
        while (!err.set && (is_punct("*") || is_punct("/"))) {
            if (!count_op()) return l;
            ExprP e(new Expr); e->k = Expr::Binary; e->op = cur().s[0]; e->at = cur(); ++pos;
            e->a = std::move(l); e->b = primary(); l = std::move(e);
        }
        return l;
    }
    ExprP primary() {
        ExprP e(new Expr); e->at = cur();
        if (is_punct("(")) {
            if (expr_depth >= kMaxExprDepth) { err.at(cur(), "expression nested deeper than 64 parentheses"); return e; }
            ++pos; ++expr_depth;
            e = additive();
            --expr_depth;
            if (err.set) return e;
            expect(")"); return postfix(std::move(e));
        }
        if (cur().kind == Tok::Num) { e->k = Expr::Literal; e->lit = cur().num; ++pos; return e; }
        if (is_punct("-") && peek(1).kind == Tok::Num) { e->k = Expr::Literal; e->lit = -peek(1).num; pos += 2; return e; }   /* `x * -1.0` */
        if (cur().kind != Tok::Ident) { err.at(cur(), "expected a connection endpoint or expression"); return e; }
        std::string path = cur().s; ++pos;
        bool multi = false;
        while (is_punct("::")) {
            ++pos; multi = true;
            if (is_punct("<")) { int depth = 0; do { if (is_punct("<")) ++depth; else if (is_punct(">")) --depth; ++pos; } while (depth > 0 && cur().kind != Tok::End); continue; }
            std::string seg; if (!ident(seg, "a path segment")) return e;
            path += "::" + seg;
        }
        if (is_punct("(")) {                                 /* free-function call: names user Rust code */
            e->k = Expr::Call; e->name = path;
            int depth = 0;
            do { if (is_punct("(")) ++depth; else if (is_punct(")")) --depth; ++pos; } while (depth > 0 && cur().kind != Tok::End);
            return e;
        }
        if (multi) { err.at(e->at, "`" + path + "` is not an endpoint (endpoints are `node.field`)"); return e; }
        e->k = Expr::Ident; e->name = path;
        return postfix(std::move(e));
    }
    ExprP postfix(ExprP e) {
        for (;;) {
            if (is_punct("[")) {
                /* `oscs[2]`: one element of a node array (ast.rs: Indexed); arrays are unrolled into scalar nodes named
                   `oscs[0]` .. `oscs[N-1]` before lowering (expand_node_arrays), so the index folds into the name */
                if (e->k != Expr::Ident || peek(1).kind != Tok::Num || !(peek(2).kind == Tok::Punct && peek(2).s == "]")) {
                    err.at(cur(), "array-indexed endpoints take the form `nodes[<integer literal>].endpoint`"); return e;
                }
                const float idx = peek(1).num;
                if (idx < 0.0f || idx != (float)(int)idx) { err.at(peek(1), "an array index must be a non-negative integer literal"); return e; }
                e->name += "[" + std::to_string((int)idx) + "]";
                pos += 3;
                continue;
            }
            if (!is_punct(".")) return e;
            if (!count_op()) return e;
            ++pos;
            ExprP f(new Expr); f->at = cur();
            if (!ident(f->name, "a field or method name")) return e;
            f->a = std::move(e);
            if (accept("(")) {
                f->k = Expr::Method;
                while (!err.set && !is_punct(")") && cur().kind != Tok::End) { float v; if (!number(v)) return f; f->args.push_back(v); if (!accept(",")) break; }
                expect(")");
            } else f->k = Expr::Field;
            e = std::move(f);
        }
    }
};

/* -------------------------------------------------------------------- node registry -- */
struct Endpoint { const char *name; Kind kind; bool is_out; int slot; /* in: stream index or value slot; out: port */ int standin; float def;
                  int width = 1;   /* lanes of a stream endpoint: `[f32; N]` fields are N wide (frame.rs) */ };
struct TypeInfo {
    const char *type;
    std::vector<Endpoint> eps;
    bool stateful;
};

/* endpoints as the reference structs declare them (#[input(..)] / #[output(..)] fields):
   oscen-lib/src/{envelope/adsr.rs:20-38, filters/tpt/mod.rs:12-24, oscillators/mod.rs:88-103, gain/mod.rs:5-12, sample_player/mod.rs:64-72},
   examples/fm-synth/src/nodes/{fm_operator.rs:12-31, crossfade.rs:8-18, mixer.rs:4-12, add_value.rs:5-13},
   examples/nih-twin-peaks/src/lp18_filter.rs:7-42, examples/oversampled-saturator/src/main.rs:31-38.
   `standin` = parameter slot read when a stream input has no stream source; `def` = the field's constructor default. */
inline const TypeInfo *type_info(const std::string &type) {
    static const std::vector<TypeInfo> reg = {
        {"AdsrEnvelope", {{"gate", Kind::Event, false, 0, -1, 0}, {"attack", Kind::Value, false, 0, -1, 0}, {"decay", Kind::Value, false, 1, -1, 0},
                          {"sustain", Kind::Value, false, 2, -1, 0}, {"release", Kind::Value, false, 3, -1, 0}, {"output", Kind::Stream, true, 0, -1, 0}}, true},
        {"FmOperator", {{"base_freq", Kind::Value, false, 0, -1, 440.0f}, {"ratio", Kind::Value, false, 1, -1, 1.0f}, {"feedback", Kind::Value, false, 2, -1, 0},
                        {"level", Kind::Value, false, 3, -1, 1.0f}, {"phase_mod", Kind::Stream, false, 0, 4, 0}, {"envelope", Kind::Stream, false, 1, 5, 1.0f},
                        {"output", Kind::Stream, true, 0, -1, 0}}, true},
        {"TptFilter", {{"q", Kind::Value, false, 0, -1, 0}, {"input", Kind::Stream, false, 0, 3, 0}, {"cutoff", Kind::Stream, false, 1, 1, 0},
                       {"f_mod", Kind::Stream, false, 2, 2, 0}, {"output", Kind::Stream, true, 0, -1, 0}}, true},
        {"PolyBlepOscillator", {{"frequency", Kind::Value, false, 0, -1, 0}, {"amplitude", Kind::Value, false, 1, -1, 0}, {"pulse_width", Kind::Value, false, 2, -1, 0.5f},
                                {"frequency_mod", Kind::Stream, false, 0, 3, 0}, {"phase_mod", Kind::Stream, false, 1, 4, 0}, {"output", Kind::Stream, true, 0, -1, 0}}, true},
        {"LP18Filter", {{"cutoff", Kind::Value, false, 0, -1, 0}, {"resonance", Kind::Value, false, 1, -1, 0}, {"fmod", Kind::Value, false, 2, -1, 0},
                        {"input", Kind::Stream, false, 0, 3, 0}, {"output", Kind::Stream, true, 0, -1, 0}}, true},
        {"IirLowpass", {{"cutoff", Kind::Value, false, 0, -1, 1000.0f}, {"q", Kind::Value, false, 1, -1, 0.70710678f}, {"input", Kind::Stream, false, 0, 2, 0},
                        {"output", Kind::Stream, true, 0, -1, 0}}, true},
        {"Oscillator", {{"frequency", Kind::Value, false, 0, -1, 0}, {"amplitude", Kind::Value, false, 1, -1, 0}, {"frequency_mod", Kind::Stream, false, 0, 2, 0},
                        {"output", Kind::Stream, true, 0, -1, 0}}, true},
        {"Delay", {{"delay_samples", Kind::Value, false, 0, -1, 0}, {"feedback", Kind::Value, false, 1, -1, 0}, {"input", Kind::Stream, false, 0, 2, 0},
                   {"output", Kind::Stream, true, 0, -1, 0}}, true},
        {"Gain", {{"input", Kind::Stream, false, 0, -1, 0}, {"gain", Kind::Stream, false, 1, 0, 1.0f}, {"output", Kind::Stream, true, 0, -1, 0}}, false},
        {"AddValue", {{"input", Kind::Stream, false, 0, -1, 0}, {"value", Kind::Value, false, 0, -1, 0}, {"output", Kind::Stream, true, 0, -1, 0}}, false},
        {"Mixer", {{"input_a", Kind::Stream, false, 0, -1, 0}, {"input_b", Kind::Stream, false, 1, -1, 0}, {"output", Kind::Stream, true, 0, -1, 0}}, false},
        {"Crossfade", {{"input", Kind::Stream, false, 0, -1, 0}, {"mix", Kind::Value, false, 0, -1, 0}, {"output_a", Kind::Stream, true, 0, -1, 0},
                       {"output_b", Kind::Stream, true, 1, -1, 0}}, false},
        {"HardClip", {{"input", Kind::Stream, false, 0, -1, 0}, {"output", Kind::Stream, true, 0, -1, 0}}, false},
        {"SamplePlayer", {{"buf", Kind::Asset, false, 0, -1, 0}, {"output", Kind::Stream, true, 0, -1, 0}}, true},
        /* examples/electric-piano/src/electric_piano_voice.rs:74-100,173-222: `amplitudes` is `[f32; NUM_HARMONICS]` on both sides */
        {"AmplitudeSource", {{"frequency", Kind::Value, false, 0, -1, 440.0f}, {"gate", Kind::Event, false, 0, -1, 0}, {"brightness", Kind::Value, false, 1, -1, 30.0f},
                             {"velocity_scaling", Kind::Value, false, 2, -1, 50.0f}, {"decay_rate", Kind::Value, false, 3, -1, 90.0f},
                             {"harmonic_decay", Kind::Value, false, 4, -1, 70.0f}, {"key_scaling", Kind::Value, false, 5, -1, 50.0f},
                             {"release_rate", Kind::Value, false, 6, -1, 40.0f}, {"amplitudes", Kind::Stream, true, 0, -1, 0, 32}}, true},
        {"OscillatorBank", {{"frequency", Kind::Value, false, 0, -1, 440.0f}, {"gate", Kind::Event, false, 0, -1, 0},
                            {"amplitudes", Kind::Stream, false, 0, -1, 0, 32}, {"output", Kind::Stream, true, 0, -1, 0}}, true},
    };
    for (const TypeInfo &t : reg) if (type == t.type) return &t;
    return nullptr;
}

/* the node library as text, one node type per line: `Type(ctor(args) | ...): endpoint kind [= default], ... -> outputs`
   (for tools: graphc --nodes, oscen_b200_node_library) */
inline std::string node_library_text() {
    static const char *const ctors[][2] = {
        {"AdsrEnvelope", "new(attack, decay, sustain, release)"}, {"FmOperator", "new()"}, {"TptFilter", "new(cutoff, q)"},
        {"PolyBlepOscillator", "sine(frequency, amplitude) | saw(..) | square(..) | triangle(..)"},
        {"LP18Filter", "new(cutoff, resonance)"}, {"IirLowpass", "new(cutoff, q)"}, {"Oscillator", "sine(frequency, amplitude) | square(..) | saw(..)"},
        {"Delay", "new(delay_samples, feedback) | from_seconds(seconds, feedback, sample_rate)"}, {"Gain", "new(gain)"}, {"AddValue", "new(value)"},
        {"Mixer", "new()"}, {"Crossfade", "new()"}, {"HardClip", "new()"}, {"SamplePlayer", "new()"}};
    std::string out;
    for (const auto &c : ctors) {
        const TypeInfo *ti = type_info(c[0]);
        if (!ti) continue;
        out += std::string(c[0]) + "::" + c[1] + "\n";
        for (const Endpoint &ep : ti->eps) {
            char b[128];
            if (ep.is_out) snprintf(b, sizeof b, "    output %s: %s\n", ep.name, "stream");
            else if (ep.kind == Kind::Event) snprintf(b, sizeof b, "    input  %s: event\n", ep.name);
            else if (ep.kind == Kind::Asset) snprintf(b, sizeof b, "    input  %s: asset (an `external NAME: AudioAsset;` slot; unconnected: silence)\n", ep.name);
            else snprintf(b, sizeof b, "    input  %s: %s (unconnected: %s)\n", ep.name, ep.kind == Kind::Value ? "value" : "stream",
                          (std::string(c[1]).find(ep.name) != std::string::npos) ? "the constructor argument" : (std::to_string(ep.def).c_str()));
            out += b;
        }
    }
    return out;
}

/* --------------------------------------------------------------------------- lowering -- */
struct Operand {
    enum K { None, Stream, Value, Const } k = None;
    plan::StreamRef s;
    uint32_t pid = plan::kNoValue;
    float c = 0.0f;
    bool cross = false;                  /* the term went through a cross-rate resampler (no fan-in summing) */
    bool fb = false;                     /* the outgoing leg of a feedback via: reads the Delay's previous-frame output */
    Kind kind() const { return k == Stream ? Kind::Stream : Kind::Value; }
};

class Lowering {
public:
    Lowering(const Ast &a, plan::Graph &out, Error &e) : ast(a), g(out), err(e) {}

    bool run() {
        g = plan::Graph();
        g.name = ast.name;
        /* ---- graph endpoints */
        for (const InputDecl &d : ast.inputs) {
            if (names.count(d.name)) { err.at(d.at, "duplicate name `" + d.name + "`"); return false; }
            names[d.name] = 'i';
            switch (d.kind) {
            case Kind::Value:
                if (g.inputs.size() >= 28) { err.at(d.at, "more than 28 value inputs (OSCEN_B200_MAX_INPUTS)"); return false; }
                value_id[d.name] = (uint32_t)g.inputs.size(); g.inputs.push_back(d.spec); break;
            case Kind::Stream:
                if (!g.stream_input.empty()) { err.at(d.at, "the node plan takes one stream input per graph (`" + g.stream_input + "` is already declared)"); return false; }
                g.stream_input = d.name; break;
            case Kind::Event:
                if (!g.event_input.empty()) { err.at(d.at, "the node plan takes one event input per graph (`" + g.event_input + "` is already declared)"); return false; }
                g.event_input = d.name; break;
            default: err.at(d.at, "asset inputs are not supported by the node plan"); return false;
            }
        }
        /* `external NAME: AudioAsset;` — a slot the host loads audio into at run time (oscen_b200_load_asset) */
        for (const ExternalDecl &d : ast.externals) {
            if (names.count(d.name)) { err.at(d.at, "duplicate name `" + d.name + "`"); return false; }
            if (d.type != "AudioAsset") { err.at(d.at, "external `" + d.name + "`: only AudioAsset slots are supported (found `" + d.type + "`)"); return false; }
            if (g.assets.size() >= 4) { err.at(d.at, "more than 4 external asset slots (OSCEN_B200_MAX_ASSETS)"); return false; }
            names[d.name] = 'x'; g.assets.push_back(d.name);
        }
        for (const OutputDecl &d : ast.outputs) {
            if (names.count(d.name)) { err.at(d.at, "duplicate name `" + d.name + "`"); return false; }
            if (d.kind != Kind::Stream) { err.at(d.at, std::string("the node plan renders stream outputs only (`") + d.name + "` is " + kind_name(d.kind) + ")"); return false; }
            if (!g.output_name.empty()) { err.at(d.at, "the node plan renders one stream output per graph (`" + g.output_name + "` is already declared)"); return false; }
            if (d.lanes != 1) { err.at(d.at, "a `Frame<2>` output needs a node that produces one: in this library that is the `Tremolo` after a polyphonic graph's voice sum"); return false; }
            names[d.name] = 'o'; g.output_name = d.name;
        }
        if (g.output_name.empty()) { Tok t; err.at(t, "graph declares no stream output"); return false; }
        /* ---- nodes */
        for (const NodeDecl &d : ast.nodes) if (!node(d)) return false;
        /* ---- connections, in statement order */
        in_terms.assign(g.nodes.size(), std::vector<std::vector<Operand>>(plan::kMaxStreamIn));
        for (const Conn &c : ast.conns) if (!connection(c)) return false;
        /* ---- fan-in: one source = plain copy, more = connect the first, accumulate the rest */
        const size_t n_decl = ast.nodes.size();              /* synthesised expression nodes follow the declared ones */
        for (size_t i = 0; i < n_decl; ++i)
            for (int k = 0; k < plan::kMaxStreamIn; ++k) if (!in_terms[i][k].empty() && !bind_stream_input((int)i, k)) return false;
        /* stateless nodes read their stream inputs unconditionally: an unconnected one holds its constructor default */
        for (size_t i = 0; i < n_decl; ++i) {
            if (info[i]->stateful) continue;
            for (const Endpoint &ep : info[i]->eps) {
                if (ep.is_out || ep.kind != Kind::Stream || ep.standin >= 0 || g.nodes[i].in[ep.slot].node != plan::kNone) continue;
                Operand c; c.k = Operand::Const; c.c = ep.def;
                const plan::StreamRef r = materialise(c, g.nodes[i].name + "." + ep.name + " (unconnected)").s;
                g.nodes[i].in[ep.slot] = r;
            }
        }
        if (out_terms.empty()) { Tok t; err.at(t, "nothing is connected to the output `" + g.output_name + "`"); return false; }
        Operand o = sum_terms(out_terms, g.output_name);
        if (o.k != Operand::Stream || o.s.node < 0) o = materialise(o, g.output_name);   /* a bare value / graph input copied to the output */
        g.output = o.s;
        if (err.set || !drop_dead_nodes()) return false;
        if (!check_rate_phases()) return false;
        /* ---- SoA state layout: nodes in declaration order */
        plan::assign_state_layout(g, 48000.0f);      /* provisional: the engine lays the state out again for its sample rate */
        return !err.set;
    }

private:
    const Ast &ast;
    plan::Graph &g;
    Error &err;
    std::map<std::string, char> names;                 /* 'i' input, 'o' output, 'n' node */
    std::map<std::string, uint32_t> value_id;
    std::map<std::string, int> node_id;
    std::vector<const TypeInfo *> info;                /* per plan node (nullptr for synthesised nodes) */
    std::vector<std::vector<std::vector<Operand>>> in_terms;   /* [node][stream input] -> sources in statement order */
    std::vector<Operand> out_terms;

    bool node(const NodeDecl &d) {
        if (names.count(d.name)) { err.at(d.at, "duplicate name `" + d.name + "`"); return false; }
        if (d.array) { err.at(d.at, "node arrays (`[T; N]`) are not compiled to a node plan: the engine's voices_per_instance is the array axis"); return false; }
        if (d.rate < 0) { err.at(d.at, "node undersampling (`/ N`) is not yet supported in v1; only oversampling (`* N`) is implemented"); return false; }
        int rate_log2 = 0;
        for (int r = d.rate; r > 1; r >>= 1) ++rate_log2;
        if (rate_log2) {
            if (g.inner_log2 && g.inner_log2 != rate_log2) { err.at(d.at, "one oversampling factor per graph: `" + d.name + "` is * " + std::to_string(d.rate) + " but an earlier node is * " + std::to_string(1 << g.inner_log2)); return false; }
            g.inner_log2 = rate_log2;
        }
        const TypeInfo *ti = type_info(d.type);
        if (!ti) { err.at(d.at, "unknown node type `" + d.type + "` (known: AdsrEnvelope, FmOperator, TptFilter, PolyBlepOscillator, Oscillator, LP18Filter, IirLowpass, Delay, SamplePlayer, Gain, AddValue, Mixer, Crossfade, HardClip, AmplitudeSource, OscillatorBank)"); return false; }
        if (!d.args_ok) { err.at(d.at, "constructor arguments of `" + d.name + "` must be numeric literals"); return false; }
        plan::Node n{};
        n.name = d.name; n.gate = false; n.rate_log2 = rate_log2;
        for (int k = 0; k < plan::kMaxParams; ++k) { n.pid[k] = plan::kNoValue; n.cst[k] = 0.0f; }
        for (const Endpoint &ep : ti->eps) if (!ep.is_out && ep.kind != Kind::Event) {
            const int slot = ep.kind == Kind::Value ? ep.slot : ep.standin;
            if (slot >= 0) n.cst[slot] = ep.def;
        }
        auto ctor = [&](const char *name, size_t n_args) {
            if (d.ctor != name) return false;
            if (d.args.size() != n_args) err.at(d.at, d.type + "::" + name + " takes " + std::to_string(n_args) + " argument(s), found " + std::to_string(d.args.size()));
            return true;
        };
        const std::string &t = d.type;
        bool ok = false;
        if (t == "AdsrEnvelope") { n.type = plan::NodeType::Adsr; if ((ok = ctor("new", 4)) && !err.set) for (int k = 0; k < 4; ++k) n.cst[k] = d.args[k]; }
        else if (t == "FmOperator") { n.type = plan::NodeType::FmOperator; ok = ctor("new", 0); }
        else if (t == "TptFilter") {
            n.type = plan::NodeType::Tpt;
            if ((ok = ctor("new", 2)) && !err.set) { n.cst[1] = d.args[0]; n.cst[0] = d.args[1]; n.init[0] = d.args[0]; n.init[1] = d.args[1]; }
        } else if (t == "PolyBlepOscillator") {
            struct { const char *c; plan::NodeType nt; } waves[] = {{"sine", plan::NodeType::PolyBlepSine}, {"saw", plan::NodeType::PolyBlepSaw},
                                                                  {"square", plan::NodeType::PolyBlepSquare}, {"triangle", plan::NodeType::PolyBlepTriangle}};
            for (auto &w : waves) if (d.ctor == w.c) { n.type = w.nt; ok = ctor(w.c, 2); }
            if (ok && !err.set) { n.cst[0] = d.args[0]; n.cst[1] = d.args[1]; }
        } else if (t == "LP18Filter") {
            n.type = plan::NodeType::Lp18;
            if ((ok = ctor("new", 2)) && !err.set) {
                float r = d.args[1]; if (r < 0.0f) r = 0.0f; if (r > 0.99f) r = 0.99f;     /* LP18Filter::new clamps the field, not last_resonance */
                n.cst[0] = d.args[0]; n.cst[1] = r; n.init[0] = d.args[0]; n.init[1] = d.args[1];
            }
        } else if (t == "IirLowpass") {
            n.type = plan::NodeType::IirLowpass;
            if ((ok = ctor("new", 2)) && !err.set) { n.cst[0] = d.args[0]; n.cst[1] = d.args[1]; n.init[0] = d.args[0]; n.init[1] = d.args[1]; }
        } else if (t == "Oscillator") {
            struct { const char *c; plan::NodeType nt; } waves[] = {{"sine", plan::NodeType::OscSine}, {"square", plan::NodeType::OscSquare}, {"saw", plan::NodeType::OscSaw}};
            for (auto &w : waves) if (d.ctor == w.c) { n.type = w.nt; ok = ctor(w.c, 2); }
            if (ok && !err.set) { n.cst[0] = d.args[0]; n.cst[1] = d.args[1]; }
        } else if (t == "Delay") {
            n.type = plan::NodeType::Delay;
            if (d.ctor == "from_seconds") {                     /* delay/mod.rs:42-45: delay_seconds * sample_rate */
                if ((ok = ctor("from_seconds", 3)) && !err.set) { n.cst[0] = d.args[0] * d.args[2]; n.cst[1] = d.args[1]; }
            } else if ((ok = ctor("new", 2)) && !err.set) { n.cst[0] = d.args[0]; n.cst[1] = d.args[1]; }
        } else if (t == "Gain") { n.type = plan::NodeType::Gain; if ((ok = ctor("new", 1)) && !err.set) n.cst[0] = d.args[0]; }
        else if (t == "AddValue") { n.type = plan::NodeType::AddValue; if ((ok = ctor("new", 1)) && !err.set) n.cst[0] = d.args[0]; }
        else if (t == "Mixer") { n.type = plan::NodeType::Mixer; ok = ctor("new", 0); }
        else if (t == "Crossfade") { n.type = plan::NodeType::Crossfade; ok = ctor("new", 0); }
        else if (t == "HardClip") { n.type = plan::NodeType::HardClip; ok = ctor("new", 0); }
        else if (t == "SamplePlayer") { n.type = plan::NodeType::SamplePlayer; ok = ctor("new", 0); }
        else if (t == "AmplitudeSource") { n.type = plan::NodeType::AmpSource; ok = ctor("new", 0); }
        else if (t == "OscillatorBank") { n.type = plan::NodeType::OscBank; ok = ctor("new", 0); }
        if (err.set) return false;
        if (!ok) { err.at(d.at, "unknown constructor `" + d.type + "::" + d.ctor + "`"); return false; }
        names[d.name] = 'n'; node_id[d.name] = (int)g.nodes.size();
        g.nodes.push_back(n); info.push_back(ti);
        return true;
    }

    int width_of(const Operand &o) const {
        return (o.k == Operand::Stream && o.s.node >= 0) ? plan::stream_width(g.nodes[o.s.node].type, o.s.port) : 1;
    }
    const Endpoint *find_ep(int node, const std::string &name) const {
        for (const Endpoint &ep : info[node]->eps) if (name == ep.name) return &ep;
        return nullptr;
    }
    int add_synth(plan::Node n) {
        g.nodes.push_back(n); info.push_back(nullptr);
        if (in_terms.size() < g.nodes.size()) in_terms.resize(g.nodes.size(), std::vector<std::vector<Operand>>(plan::kMaxStreamIn));
        return (int)g.nodes.size() - 1;
    }
    plan::Node blank(const std::string &name, plan::NodeType t, plan::ExprOp op) {
        plan::Node n{};
        n.name = name; n.type = t; n.op = op; n.gate = false;
        for (int k = 0; k < plan::kMaxParams; ++k) { n.pid[k] = plan::kNoValue; n.cst[k] = 0.0f; }
        return n;
    }
    static Operand stream_of(int node, int port = 0) { Operand o; o.k = Operand::Stream; o.s.node = node; o.s.port = port; return o; }
    /* log2 of the rate a stream is produced at (0 = the graph's rate); -1 for values and literals, which fit any rate */
    int rate_of(const Operand &o) const {
        if (o.k != Operand::Stream) return -1;
        if (o.s.node < 0) return 0;
        const plan::Node &n = g.nodes[o.s.node];
        return n.type == plan::NodeType::Down ? 0 : n.rate_log2;
    }

    /* a value (graph input or literal) read as a stream: one broadcast per frame */
    Operand materialise(const Operand &v, const std::string &label) {
        if (v.k == Operand::Stream && v.s.node >= 0) return v;
        plan::Node n = blank(label, plan::NodeType::Expr, plan::ExprOp::Value);
        if (v.k == Operand::Stream) {                        /* the graph's stream input copied through: x + 0 would change -0.0, so copy = x * 1 */
            n.op = plan::ExprOp::MulSV; n.in[0] = v.s; n.fb[0] = v.fb; n.cst[0] = 1.0f;
        } else if (v.k == Operand::Value) n.pid[0] = v.pid; else n.cst[0] = v.c;
        return stream_of(add_synth(n));
    }
    Operand binary(char op, const Operand &l, const Operand &r, const std::string &label, const Tok &at) {
        using O = plan::ExprOp;
        const bool ls = l.k == Operand::Stream, rs = r.k == Operand::Stream;
        if (!ls && !rs) {
            if (l.k == Operand::Const && r.k == Operand::Const) {     /* literal (op) literal: one f32 operation, folded */
                Operand o; o.k = Operand::Const;
                o.c = op == '+' ? l.c + r.c : op == '-' ? l.c - r.c : op == '*' ? l.c * r.c : l.c / r.c;
                return o;
            }
            return binary(op, materialise(l, label + " (lhs)"), r, label, at);
        }
        plan::Node n = blank(label, plan::NodeType::Expr, O::None);
        if (ls && rs && rate_of(l) != rate_of(r)) {
            err.at(at, "v1 does not support arithmetic between streams of different rates; route through an outer-rate node instead");
            return l;
        }
        n.rate_log2 = ls ? rate_of(l) : rate_of(r);
        if (ls && rs) {
            n.op = op == '+' ? O::AddSS : op == '-' ? O::SubSS : op == '*' ? O::MulSS : O::DivSS;
            n.in[0] = l.s; n.in[1] = r.s; n.fb[0] = l.fb; n.fb[1] = r.fb;
        } else {
            const Operand &s = ls ? l : r, &v = ls ? r : l;
            /* a + v == v + a and a * v == v * a bit for bit; - and / keep their operand order */
            if (op == '+') n.op = O::AddSV; else if (op == '*') n.op = O::MulSV;
            else if (op == '-') n.op = ls ? O::SubSV : O::SubVS; else n.op = ls ? O::DivSV : O::DivVS;
            n.in[0] = s.s; n.fb[0] = s.fb;
            if (v.k == Operand::Value) n.pid[0] = v.pid; else n.cst[0] = v.c;
        }
        (void)at;
        return stream_of(add_synth(n));
    }
    std::string text_of(const Expr &e) const {
        switch (e.k) {
        case Expr::Ident: return e.name;
        case Expr::Field: return text_of(*e.a) + "." + e.name;
        case Expr::Method: { std::string s = text_of(*e.a) + "." + e.name + "(";
            for (size_t i = 0; i < e.args.size(); ++i) { if (i) s += ", "; char b[32]; snprintf(b, sizeof b, "%g", (double)e.args[i]); s += b; }
            return s + ")"; }
        case Expr::Binary: return text_of(*e.a) + " " + e.op + " " + text_of(*e.b);
        case Expr::Literal: { char b[32]; snprintf(b, sizeof b, "%g", (double)e.lit); return b; }
        default: return e.name + "(..)";
        }
    }
    /* source expression -> operand (lower.rs:1182-: lower_expr; kinds per endpoint_kind_of, lower.rs:937-959) */
    bool source(const Expr &e, Operand &out, bool &is_event) {
        is_event = false;
        switch (e.k) {
        case Expr::Literal: out.k = Operand::Const; out.c = e.lit; return true;
        case Expr::Ident: {
            auto it = names.find(e.name);
            if (it == names.end()) { err.at(e.at, "cannot resolve the source of this connection (unknown node or endpoint name `" + e.name + "`?)"); return false; }
            if (it->second == 'o') { err.at(e.at, "graph output `" + e.name + "` cannot be a connection source"); return false; }
            if (it->second == 'n') { err.at(e.at, "`" + e.name + "` is a node: name one of its outputs (`" + e.name + ".output`)"); return false; }
            if (it->second == 'x') { err.at(e.at, "`" + e.name + "` is an external asset slot: it connects to an asset endpoint only (`" + e.name + " -> player.buf`)"); return false; }
            if (e.name == g.stream_input) { out.k = Operand::Stream; out.s.node = plan::kGraphInput; out.s.port = 0; return true; }
            if (e.name == g.event_input) { is_event = true; return true; }
            out.k = Operand::Value; out.pid = value_id[e.name]; return true;
        }
        case Expr::Field: {
            if (e.a->k != Expr::Ident || !node_id.count(e.a->name)) {
                err.at(e.at, "cannot resolve the source of this connection (unknown node or endpoint name `" + text_of(*e.a) + "`?)"); return false;
            }
            const int nid = node_id[e.a->name];
            const Endpoint *ep = find_ep(nid, e.name);
            if (!ep) { err.at(e.at, "node `" + e.a->name + "` has no endpoint `" + e.name + "`"); return false; }
            if (!ep->is_out) { err.at(e.at, "`" + e.a->name + "." + e.name + "` is an input: it cannot be a connection source"); return false; }
            out = stream_of(nid, ep->slot); return true;
        }
        case Expr::Binary: {
            Operand l, r; bool le, re;
            if (!source(*e.a, l, le) || !source(*e.b, r, re)) return false;
            if (le || re) { err.at(e.at, "event endpoints cannot take part in arithmetic"); return false; }
            if (width_of(l) > 1 || width_of(r) > 1) { err.at(e.at, "an array-typed stream (`[f32; N]`) connects one output to one input directly: no arithmetic on it"); return false; }
            out = binary(e.op, l, r, text_of(e), e.at); return !err.set;
        }
        case Expr::Method: {
            using O = plan::ExprOp;
            Operand x; bool xe;
            if (!source(*e.a, x, xe)) return false;
            if (xe) { err.at(e.at, "event endpoints cannot take part in arithmetic"); return false; }
            if (width_of(x) > 1) { err.at(e.at, "an array-typed stream (`[f32; N]`) connects one output to one input directly: no arithmetic on it"); return false; }
            struct { const char *m; O op; size_t n_args; } ms[] = {{"tanh", O::Tanh, 0}, {"abs", O::Abs, 0}, {"clamp", O::Clamp, 2}, {"min", O::MinV, 1}, {"max", O::MaxV, 1},
                                                                   {"sqrt", O::Sqrt, 0}, {"sin", O::Sin, 0}, {"cos", O::Cos, 0}, {"exp", O::Exp, 0}};
            for (auto &m : ms) if (e.name == m.m) {
                if (e.args.size() != m.n_args) { err.at(e.at, "f32::" + e.name + " takes " + std::to_string(m.n_args) + " literal argument(s)"); return false; }
                plan::Node n = blank(text_of(e), plan::NodeType::Expr, m.op);
                const Operand xs = materialise(x, text_of(*e.a));
                n.in[0] = xs.s; n.fb[0] = xs.fb; n.rate_log2 = rate_of(xs);
                for (size_t i = 0; i < e.args.size(); ++i) n.cst[i] = e.args[i];
                out = stream_of(add_synth(n)); return true;
            }
            err.at(e.at, "method `." + e.name + "()` is not supported in a connection source (tanh, abs, clamp, min, max, sqrt, sin, cos, exp are)"); return false;
        }
        default:
            err.at(e.at, "free function `" + e.name + "` in a connection source names Rust code: not supported by the node plan"); return false;
        }
    }

    bool connection(const Conn &c) {
        /* `sample -> player.buf`: an external asset slot bound to a node's asset endpoint (not a per-frame edge: the node
           reads whatever buffer the slot holds, sample_player/mod.rs:104-112) */
        if (c.src->k == Expr::Ident && names.count(c.src->name) && names[c.src->name] == 'x') {
            const Expr &d = *c.dst;
            if (c.via) { err.at(c.at, "a `-> [..] ->` edge carries a stream"); return false; }
            if (d.k != Expr::Field || d.a->k != Expr::Ident || !node_id.count(d.a->name)) { err.at(d.at, "an external asset connects to a node's asset endpoint (`" + c.src->name + " -> player.buf`)"); return false; }
            const int nid = node_id[d.a->name];
            const Endpoint *ep = find_ep(nid, d.name);
            if (!ep) { err.at(d.at, "node `" + d.a->name + "` has no endpoint `" + d.name + "`"); return false; }
            if (ep->is_out || ep->kind != Kind::Asset) { err.at(c.at, std::string("Type mismatch in connection: source is Asset but destination expects ") + kind_name(ep->kind)); return false; }
            for (size_t k = 0; k < g.assets.size(); ++k) if (g.assets[k] == c.src->name) g.nodes[nid].asset = (int)k;
            return true;
        }
        Operand src; bool src_event = false;
        if (!source(*c.src, src, src_event)) return false;
        if (c.via) {
            /* `src -> [N] -> dst` routes the edge through a hidden Delay::new(N as f32, 0.0), `src -> [name] -> dst` through
               a declared Delay (ir/lower.rs:560-656): src feeds the delay's input in the normal order; the leg out of the
               delay does not order the walk, so dst reads what the delay output when it last ran */
            if (src_event || src.k != Operand::Stream) { err.at(c.at, "a `-> [..] ->` edge carries a stream"); return false; }
            if (c.dst->k != Expr::Field) { err.at(c.at, "a `-> [..] ->` edge must end at a node input"); return false; }
            int did;
            if (c.via == 1) {
                plan::Node n = blank(text_of(*c.src) + " -> [" + std::to_string((long)c.via_samples) + "]", plan::NodeType::Delay, plan::ExprOp::None);
                n.cst[0] = c.via_samples; n.cst[1] = 0.0f;
                n.in[0] = src.s; n.fb[0] = src.fb; n.rate_log2 = rate_of(src) > 0 ? rate_of(src) : 0;
                did = add_synth(n);
            } else {
                if (!node_id.count(c.via_node)) { err.at(c.at, "`-> [" + c.via_node + "] ->`: no such node"); return false; }
                did = node_id[c.via_node];
                if (g.nodes[did].type != plan::NodeType::Delay) {
                    err.at(c.at, "`-> [" + c.via_node + "] ->`: `" + c.via_node + "` does not implement AllowsFeedback (only Delay does)"); return false;
                }
                if (rate_of(src) > 0 && rate_of(src) != g.nodes[did].rate_log2) { err.at(c.at, "a feedback edge cannot cross a rate boundary"); return false; }
                in_terms[did][0].push_back(src);
            }
            src = stream_of(did);
            src.fb = true;
        }
        const Kind sk = src_event ? Kind::Event : src.kind();
        const Expr &d = *c.dst;
        /* array-typed streams (`[f32; N]`): a plain edge between two endpoints of the same width — no arithmetic, no fan-in,
           no resampler, no delay (the reference's Frame arithmetic is element-wise; nothing in its graphs uses it on arrays) */
        const int sw = width_of(src);
        auto width_name = [](int w) { return w == 1 ? std::string("f32") : "[f32; " + std::to_string(w) + "]"; };
        if (sw > 1 && (c.src->k != Expr::Field || c.via)) { err.at(c.at, "an array-typed stream (`" + width_name(sw) + "`) connects one output to one input directly"); return false; }
        auto mismatch = [&](Kind dk) {
            err.at(c.at, std::string("Type mismatch in connection: source is ") + kind_name(sk) + " but destination expects " + kind_name(dk));
            return false;
        };
        if (d.k == Expr::Ident) {
            auto it = names.find(d.name);
            if (it == names.end() || it->second != 'o') {
                err.at(d.at, it == names.end() ? "cannot resolve the destination of this connection (unknown node or endpoint name `" + d.name + "`?)"
                                               : "`" + d.name + "` is not a graph output: it cannot be a connection destination");
                return false;
            }
            if (sk == Kind::Event) return mismatch(Kind::Stream);
            if (src.fb) { err.at(c.at, "a `-> [..] ->` edge must end at a node input"); return false; }
            if (sw > 1) { err.at(c.at, "Type mismatch in connection: source is " + width_name(sw) + " but the graph output is f32"); return false; }
            if (!cross_rate(c, src, 0, d.name)) return false;
            out_terms.push_back(src);
            return true;
        }
        if (d.k != Expr::Field || d.a->k != Expr::Ident || !node_id.count(d.a->name)) {
            err.at(d.at, "cannot resolve the destination of this connection (unknown node or endpoint name `" + text_of(d) + "`?)"); return false;
        }
        const int nid = node_id[d.a->name];
        const Endpoint *ep = find_ep(nid, d.name);
        if (!ep) { err.at(d.at, "node `" + d.a->name + "` has no endpoint `" + d.name + "`"); return false; }
        if (ep->is_out) { err.at(d.at, "`" + d.a->name + "." + d.name + "` is an output: it cannot be a connection destination"); return false; }
        plan::Node &n = g.nodes[nid];
        switch (ep->kind) {
        case Kind::Asset: return mismatch(Kind::Asset);
        case Kind::Event:
            if (sk != Kind::Event) return mismatch(Kind::Event);
            n.gate = true; return true;
        case Kind::Value:
            if (sk != Kind::Value) return mismatch(Kind::Value);
            if (src.k == Operand::Stream) { err.at(c.at, "a computed value into a value endpoint is not supported by the node plan (connect a graph input or a literal)"); return false; }
            /* plain field assignment every frame: the last statement wins */
            if (src.k == Operand::Value) { n.pid[ep->slot] = src.pid; } else { n.pid[ep->slot] = plan::kNoValue; n.cst[ep->slot] = src.c; }
            return true;
        default:
            if (sk == Kind::Event) return mismatch(Kind::Stream);
            if (sw != ep->width) {
                if (sk != Kind::Stream) return mismatch(Kind::Stream);
                err.at(c.at, "Type mismatch in connection: source is " + width_name(sw) + " but `" + d.a->name + "." + d.name + "` expects " + width_name(ep->width)); return false;
            }
            if (sw > 1) {
                if (!in_terms[nid][ep->slot].empty()) { err.at(c.at, "fan-in summing supports only scalar stream sources; `" + d.a->name + "." + d.name + "` is " + width_name(sw)); return false; }
                if (rate_of(src) != g.nodes[nid].rate_log2) { err.at(c.at, "an array-typed stream cannot cross a rate boundary"); return false; }
            }
            if (!cross_rate(c, src, g.nodes[nid].rate_log2, d.a->name + "." + d.name)) return false;
            in_terms[nid][ep->slot].push_back(src);
            return true;
        }
    }

    /* a stream edge whose source and destination run at different rates goes through the connection
       policy's resampler (ir/lower.rs:741-906: classify_edge_ir; stream edges default to [sinc],
       value edges latch = plain reads).  One resampler per edge, as the reference keeps one field per edge. */
    bool cross_rate(const Conn &c, Operand &src, int dst_rate, const std::string &dst_label) {
        const int sr = rate_of(src);
        if (sr < 0 || sr == dst_rate) return true;
        if (sr > 0 && dst_rate > 0) {
            err.at(c.at, "v1 does not support connections between two differently-rated non-default-rate nodes; route through an outer-rate node instead");
            return false;
        }
        const bool up = sr == 0;
        const int l = up ? dst_rate : sr;
        const int policy = c.policy == "sinc_iir" ? plan::kPolicySincIir : c.policy == "linear" ? plan::kPolicyLinear
                         : c.policy == "latch" ? plan::kPolicyLatch : plan::kPolicySinc;
        static const char *const pn[] = {"sinc", "sinc_iir", "linear", "latch"};
        plan::Node n = blank(text_of(*c.src) + " -> " + dst_label + " [" + pn[policy] + (up ? " up x" : " down x") + std::to_string(1 << l) + "]",
                             up ? plan::NodeType::Up : plan::NodeType::Down, plan::ExprOp::None);
        if (src.fb) { err.at(c.at, "a feedback edge cannot cross a rate boundary"); return false; }
        n.in[0] = src.s; n.rate_log2 = l; n.policy = policy;
        src = stream_of(add_synth(n));
        src.cross = true;
        return true;
    }

    /* codegen/emit_node.rs:153-174: dst = s0; dst += s1; ... (left fold in statement order) */
    Operand sum_terms(const std::vector<Operand> &terms, const std::string &label) {
        if (terms.size() > 1) for (const Operand &t : terms) if (t.cross) {
            Tok at;
            err.at(at, "fan-in summing supports only same-rate scalar/frame stream sources; saw a cross-rate edge into `" + label + "`");
            return terms[0];
        }
        Operand acc = terms[0];
        for (size_t i = 1; i < terms.size(); ++i) {
            Tok t;
            acc = binary('+', acc, terms[i], label + " (fan-in)", t);
            if (acc.k == Operand::Stream && acc.s.node >= 0 && g.nodes[acc.s.node].op == plan::ExprOp::AddSS) {
                g.nodes[acc.s.node].type = plan::NodeType::Mixer; g.nodes[acc.s.node].op = plan::ExprOp::None;
            }
        }
        return acc;
    }
    bool bind_stream_input(int nid, int k) {
        const Endpoint *ep = nullptr;
        for (const Endpoint &e : info[nid]->eps) if (!e.is_out && e.kind == Kind::Stream && e.slot == k) ep = &e;
        const std::string label = g.nodes[nid].name + "." + ep->name;
        const Operand o = sum_terms(in_terms[nid][k], label);
        if (o.k == Operand::Stream) { g.nodes[nid].in[k] = o.s; g.nodes[nid].fb[k] = o.fb; return true; }
        if (ep->standin >= 0) {                              /* Value -> Stream edge: the kernel broadcasts the value per frame */
            plan::Node &n = g.nodes[nid];
            if (o.k == Operand::Value) n.pid[ep->standin] = o.pid; else { n.pid[ep->standin] = plan::kNoValue; n.cst[ep->standin] = o.c; }
            return true;
        }
        g.nodes[nid].in[k] = materialise(o, label).s;
        return true;
    }

    /* multirate frame body (codegen/emit_frame.rs:114-176): outer-rate nodes that feed the oversampled section
       run before the inner loop, those that depend on a Down edge after it — a node cannot be both */
    bool check_rate_phases() {
        const int n = (int)g.nodes.size();
        g.inner_log2 = 0;
        for (const plan::Node &nd : g.nodes) if (nd.rate_log2 > 0) g.inner_log2 = nd.rate_log2;
        if (!g.inner_log2) return true;
        std::vector<char> tainted(n, 0);                 /* depends (transitively) on a Down edge */
        bool moved = true;
        while (moved) {
            moved = false;
            for (int i = 0; i < n; ++i) {
                if (tainted[i]) continue;
                bool t = g.nodes[i].type == plan::NodeType::Down;
                for (const plan::StreamRef &r : g.nodes[i].in) if (r.node >= 0 && tainted[r.node]) t = true;
                if (t) { tainted[i] = 1; moved = true; }
            }
        }
        for (int i = 0; i < n; ++i) if (g.nodes[i].type == plan::NodeType::Up && tainted[i]) {
            Tok at;
            err.at(at, "`" + g.nodes[i].name + "`: an outer-rate signal computed from the oversampled section cannot feed it again in the same frame");
            return false;
        }
        return true;
    }

    /* ir/passes/dead_nodes.rs: reverse BFS from the output; everything else is dropped */
    bool drop_dead_nodes() {
        const int n = (int)g.nodes.size();
        std::vector<char> live(n, 0);
        std::vector<int> stack{g.output.node};
        while (!stack.empty()) {
            const int i = stack.back(); stack.pop_back();
            if (i < 0 || live[i]) continue;
            live[i] = 1;
            for (const plan::StreamRef &r : g.nodes[i].in) if (r.node >= 0) stack.push_back(r.node);
        }
        std::vector<int> remap(n, -1);
        std::vector<plan::Node> kept;
        for (int i = 0; i < n; ++i) if (live[i]) { remap[i] = (int)kept.size(); kept.push_back(g.nodes[i]); }
        for (plan::Node &nd : kept) for (plan::StreamRef &r : nd.in) if (r.node >= 0) r.node = remap[r.node];
        g.output.node = remap[g.output.node];
        g.nodes.swap(kept);
        return true;
    }
};

/* ------------------------------------------------------------------ nested graphs -- */
/* A graph used as a scalar node of another graph (`voice1 = SimpleVoice;`, oscen-lib/tests/nested_graph_test.rs)
 * is just another node whose process() the parent calls: the reference inlines the child's generated walk.
 * Here the child is inlined at the AST level: its nodes join the parent as `<instance>.<node>`, its inputs are
 * replaced by what the parent connects to them (an unconnected input keeps its declared default), and a parent
 * source `<instance>.<output>` becomes the child's expression for that output (several -> their sum in statement
 * order, like any fan-in).  Values stay values, streams stay streams: the kind rules are checked afterwards on
 * the flattened graph exactly as for a hand-flattened one. */
inline ExprP clone_expr(const Expr &e) {
    ExprP c(new Expr);
    c->k = e.k; c->name = e.name; c->op = e.op; c->lit = e.lit; c->args = e.args; c->at = e.at;
    if (e.a) c->a = clone_expr(*e.a);
    if (e.b) c->b = clone_expr(*e.b);
    return c;
}
inline const Ast *find_graph(const std::vector<Ast> &all, const std::string &name);

/* Node arrays of library types — `oscs = [PolyBlepOscillator::saw(440.0, 0.6); 4]` (oscen-lib/tests/array_test.rs,
   array_fanin_stream_input.rs, multirate_array_fanout.rs) — are unrolled into scalar nodes `oscs[0]` .. `oscs[N-1]`, and
   every connection that names the array without an index takes the shape ir/graph.rs:48-78 gives it:
     scalar -> array.ep   Broadcast: one edge per element (a cross-rate broadcast gets one resampler per element here where
                          the reference shares one — same input, same state, same values)
     array.ep -> scalar   FanIn: `let mut sum = 0.0; for i { sum += elt }` (codegen/emit_edge.rs:208-233), summed at the
                          source rate, then ONE resampler if the edge crosses rates
     array.ep -> array.ep Parallel: element k to element k, for k < min(N, M)
   Arrays of GRAPHS stay with lower_polyphonic (the engine's voices_per_instance axis). */
struct ArrayExpander {
    Error &err;
    std::map<std::string, int> arrays;
    static ExprP field_of(const std::string &node, const std::string &field, const Tok &at) {
        ExprP id(new Expr); id->k = Expr::Ident; id->name = node; id->at = at;
        ExprP f(new Expr); f->k = Expr::Field; f->name = field; f->at = at; f->a = std::move(id);
        return f;
    }
    /* N if `e` is `array.endpoint` without an index, else 0 */
    int array_endpoint(const Expr &e) const {
        if (e.k != Expr::Field || !e.a || e.a->k != Expr::Ident) return 0;
        auto it = arrays.find(e.a->name);
        return it == arrays.end() ? 0 : it->second;
    }
    bool mentions_array(const Expr &e) const {
        if (e.k == Expr::Ident && arrays.count(e.name)) return true;
        if (array_endpoint(e)) return true;
        return (e.a && mentions_array(*e.a)) || (e.b && mentions_array(*e.b));
    }
    static Conn like(const Conn &c) {
        Conn k; k.policy = c.policy; k.via = c.via; k.via_samples = c.via_samples; k.via_node = c.via_node; k.at = c.at;
        return k;
    }
    bool run(Ast &g) {
        arrays.clear();
        std::vector<NodeDecl> nodes;
        for (const NodeDecl &d : g.nodes) {
            if (d.array > 0 && type_info(d.type)) {
                if (d.array > 64) { err.at(d.at, "node array `" + d.name + "`: more than 64 elements"); return false; }
                if (arrays.count(d.name)) { err.at(d.at, "duplicate name `" + d.name + "`"); return false; }
                arrays[d.name] = d.array;
                for (int k = 0; k < d.array; ++k) { NodeDecl e = d; e.array = 0; e.name = d.name + "[" + std::to_string(k) + "]"; nodes.push_back(e); }
            } else nodes.push_back(d);
        }
        if (arrays.empty()) return true;
        g.nodes.swap(nodes);
        std::vector<Conn> conns;
        for (Conn &c : g.conns) {
            const int sn = array_endpoint(*c.src), dn = array_endpoint(*c.dst);
            if (!sn && mentions_array(*c.src)) {
                err.at(c.at, "a node array inside a source expression: name its elements (`oscs[0].output + oscs[1].output`) or connect `array.endpoint` on its own"); return false;
            }
            if (!dn && mentions_array(*c.dst)) { err.at(c.at, "a node array is connected through one of its endpoints (`array.endpoint`)"); return false; }
            if ((sn || dn) && c.via) { err.at(c.at, "feedback edges (`-> [..] ->`) take scalar endpoints: index the array elements"); return false; }
            if (!sn && !dn) { conns.push_back(std::move(c)); continue; }
            if (dn && !sn) {                                          /* Broadcast */
                for (int k = 0; k < dn; ++k) {
                    Conn e = like(c);
                    e.src = clone_expr(*c.src);
                    e.dst = field_of(c.dst->a->name + "[" + std::to_string(k) + "]", c.dst->name, c.dst->at);
                    conns.push_back(std::move(e));
                }
            } else if (dn && sn) {                                    /* Parallel */
                for (int k = 0; k < std::min(sn, dn); ++k) {
                    Conn e = like(c);
                    e.src = field_of(c.src->a->name + "[" + std::to_string(k) + "]", c.src->name, c.src->at);
                    e.dst = field_of(c.dst->a->name + "[" + std::to_string(k) + "]", c.dst->name, c.dst->at);
                    conns.push_back(std::move(e));
                }
            } else {                                                  /* FanIn: 0.0 + e0 + e1 + ... */
                ExprP acc(new Expr); acc->k = Expr::Literal; acc->lit = 0.0f; acc->at = c.src->at;
                for (int k = 0; k < sn; ++k) {
                    ExprP b(new Expr); b->k = Expr::Binary; b->op = '+'; b->at = c.src->at;
                    b->a = std::move(acc);
                    b->b = field_of(c.src->a->name + "[" + std::to_string(k) + "]", c.src->name, c.src->at);
                    acc = std::move(b);
                }
                Conn e = like(c);
                e.src = std::move(acc); e.dst = clone_expr(*c.dst);
                conns.push_back(std::move(e));
            }
        }
        g.conns.swap(conns);
        return true;
    }
};

inline const Ast *find_graph(const std::vector<Ast> &all, const std::string &name) {
    for (const Ast &a : all) if (a.name == name) return &a;
    return nullptr;
}
inline ExprP sum_exprs(std::vector<ExprP> &terms) {
    ExprP acc = std::move(terms[0]);
    for (size_t i = 1; i < terms.size(); ++i) {
        ExprP b(new Expr); b->k = Expr::Binary; b->op = '+'; b->at = acc->at;
        b->a = std::move(acc); b->b = std::move(terms[i]); acc = std::move(b);
    }
    return acc;
}
struct Inliner {
    const std::vector<Ast> &all;
    Error &err;
    int depth = 0;

    /* child source expression -> parent terms: `input` -> the parent's expression for it, `node.x` -> `inst.node.x` */
    ExprP rewrite_child_source(const Expr &e, const std::string &inst, const Ast &child,
                               std::map<std::string, std::vector<ExprP>> &bound, bool &dropped) {
        if (e.k == Expr::Ident) {
            for (const InputDecl &in : child.inputs) if (in.name == e.name) {
                auto it = bound.find(e.name);
                if (it == bound.end() || it->second.empty()) {
                    if (in.kind == Kind::Event) { dropped = true; return nullptr; }       /* an unconnected event input never fires */
                    ExprP l(new Expr); l->k = Expr::Literal; l->lit = in.kind == Kind::Value ? in.def : 0.0f; l->at = e.at;
                    return l;
                }
                std::vector<ExprP> terms;
                if (in.kind == Kind::Stream) for (const ExprP &t : it->second) terms.push_back(clone_expr(*t));   /* fan-in sums */
                else terms.push_back(clone_expr(*it->second.back()));                                                /* last write wins */
                return sum_exprs(terms);
            }
            return clone_expr(e);
        }
        ExprP c(new Expr);
        c->k = e.k; c->name = e.name; c->op = e.op; c->lit = e.lit; c->args = e.args; c->at = e.at;
        if (e.k == Expr::Field && e.a && e.a->k == Expr::Ident) {                 /* node.endpoint */
            c->a.reset(new Expr); c->a->k = Expr::Ident; c->a->name = inst + "." + e.a->name; c->a->at = e.a->at;
            return c;
        }
        if (e.a) { c->a = rewrite_child_source(*e.a, inst, child, bound, dropped); if (!c->a) return nullptr; }
        if (e.b) { c->b = rewrite_child_source(*e.b, inst, child, bound, dropped); if (!c->b) return nullptr; }
        return c;
    }
    /* parent source: replace `inst.output` by the child's expression(s) for that output */
    ExprP rewrite_parent_source(const Expr &e, const std::string &inst, std::map<std::string, std::vector<ExprP>> &outs, const Ast &child) {
        if (e.k == Expr::Field && e.a && e.a->k == Expr::Ident && e.a->name == inst) {
            auto it = outs.find(e.name);
            if (it == outs.end() || it->second.empty()) {
                bool declared = false;
                for (const OutputDecl &o : child.outputs) declared |= o.name == e.name;
                err.at(e.at, declared ? "output `" + e.name + "` of `" + child.name + "` has nothing connected to it"
                                      : "graph `" + child.name + "` has no output `" + e.name + "`");
                return nullptr;
            }
            std::vector<ExprP> terms;
            for (const ExprP &t : it->second) terms.push_back(clone_expr(*t));
            return sum_exprs(terms);
        }
        ExprP c(new Expr);
        c->k = e.k; c->name = e.name; c->op = e.op; c->lit = e.lit; c->args = e.args; c->at = e.at;
        if (e.a) { c->a = rewrite_parent_source(*e.a, inst, outs, child); if (!c->a) return nullptr; }
        if (e.b) { c->b = rewrite_parent_source(*e.b, inst, outs, child); if (!c->b) return nullptr; }
        return c;
    }

    /* returns a copy of `g` with every scalar nested-graph node inlined (recursively) */
    bool flatten(const Ast &g, Ast &out) {
        if (++depth > 16) { Tok t; err.at(t, "nested graphs: recursion deeper than 16 levels (a graph that contains itself?)"); return false; }
        out.name = g.name;
        for (const ExternalDecl &d : g.externals) out.externals.push_back(d);
        for (const InputDecl &d : g.inputs) out.inputs.push_back(d);
        for (const OutputDecl &d : g.outputs) out.outputs.push_back(d);
        std::vector<const NodeDecl *> nested;
        for (const NodeDecl &d : g.nodes) {
            const Ast *child = d.array ? nullptr : find_graph(all, d.type);
            if (child && child != &g) nested.push_back(&d); else out.nodes.push_back(d);
        }
        std::vector<Conn> conns;
        for (const Conn &c : g.conns) {
            Conn k; k.policy = c.policy; k.via = c.via; k.via_samples = c.via_samples; k.via_node = c.via_node; k.at = c.at;
            k.src = clone_expr(*c.src); k.dst = clone_expr(*c.dst);
            conns.push_back(std::move(k));
        }
        for (const NodeDecl *nd : nested) {
            if (nd->rate) { err.at(nd->at, "a rate on a nested graph instance is not supported (put `* N` on the nodes inside it)"); return false; }
            if (nd->ctor != "new" || !nd->args.empty()) { err.at(nd->at, "a graph is instantiated with `" + nd->type + "` or `" + nd->type + "::new()`"); return false; }
            Ast child;
            { Inliner sub{all, err, depth}; if (!sub.flatten(*find_graph(all, nd->type), child)) return false; }
            if (!child.externals.empty()) { err.at(nd->at, "nested graph `" + nd->type + "` declares an `external` asset slot: declare it in the top-level graph"); return false; }
            const std::string &inst = nd->name;
            /* what the parent connects INTO the instance, per child input (statement order) */
            std::map<std::string, std::vector<ExprP>> bound;
            std::vector<Conn> rest;
            for (Conn &c : conns) {
                const Expr &d = *c.dst;
                if (d.k == Expr::Field && d.a->k == Expr::Ident && d.a->name == inst) {
                    bool is_input = false;
                    for (const InputDecl &in : child.inputs) is_input |= in.name == d.name;
                    if (!is_input) { err.at(d.at, "graph `" + child.name + "` has no input `" + d.name + "`"); return false; }
                    if (c.via) { err.at(c.at, "a `-> [..] ->` edge into a nested graph instance is not supported"); return false; }
                    bound[d.name].push_back(std::move(c.src));
                } else rest.push_back(std::move(c));
            }
            conns.swap(rest);
            /* the child's nodes and connections join the parent under `inst.` */
            for (NodeDecl cn : child.nodes) { cn.name = inst + "." + cn.name; out.nodes.push_back(cn); }
            std::map<std::string, std::vector<ExprP>> outs;
            for (const Conn &cc : child.conns) {
                bool dropped = false;
                ExprP src = rewrite_child_source(*cc.src, inst, child, bound, dropped);
                if (dropped) continue;
                if (!src) return false;
                if (cc.dst->k == Expr::Ident) {                                       /* -> child output */
                    bool is_out = false;
                    for (const OutputDecl &o : child.outputs) is_out |= o.name == cc.dst->name;
                    if (!is_out) { err.at(cc.dst->at, "cannot resolve the destination of this connection (unknown node or endpoint name `" + cc.dst->name + "`?)"); return false; }
                    if (cc.via) { err.at(cc.at, "a `-> [..] ->` edge must end at a node input"); return false; }
                    outs[cc.dst->name].push_back(std::move(src));
                    continue;
                }
                Conn k; k.policy = cc.policy; k.via = cc.via; k.via_samples = cc.via_samples; k.at = cc.at;
                k.via_node = cc.via == 2 ? inst + "." + cc.via_node : cc.via_node;
                k.src = std::move(src);
                k.dst = clone_expr(*cc.dst);
                if (k.dst->k == Expr::Field && k.dst->a->k == Expr::Ident) k.dst->a->name = inst + "." + k.dst->a->name;
                conns.push_back(std::move(k));
            }
            /* parent sources that read the instance's outputs */
            for (Conn &c : conns) {
                ExprP src = rewrite_parent_source(*c.src, inst, outs, child);
                if (!src) return false;
                c.src = std::move(src);
            }
        }
        out.conns = std::move(conns);
        --depth;
        return true;
    }
};

/* ------------------------------------------------------------- polyphonic wrapper -- */
/* An outer graph of the shape the reference's polyphonic examples have (examples/fm-synth/src/lib.rs:22-131):
 *     nodes { midi_parser = MidiParser::new(); voice_allocator = VoiceAllocator::<N>::new();
 *             voice_handlers = [MidiVoiceHandler::new(); N]; voices = [Voice::new(); N]; }
 * where `Voice` is another graph of the same text.  The reference expands it into N copies of the voice
 * struct walked in index order (codegen/emit_node.rs:365-379) and sums `voices.<out>` in index order
 * (:463-466).  Here the voice graph is what gets planned; the array axis becomes the engine's
 * voices_per_instance, broadcast edges `outer_value -> voices.<input>` become names for the voice's inputs
 * (with the OUTER declaration's default and ramp), and the MIDI chain selects the engine's MIDI front
 * (oscen_b200_push_midi / _push_midi_batch: midi.rs:69-224, voice_allocator.rs:57-136). */
inline bool is_polyphonic(const Ast &top) {
    for (const NodeDecl &d : top.nodes) if (d.array > 0) return true;
    return false;
}
inline bool lower_polyphonic(const Ast &top, const std::vector<Ast> &all, plan::Graph &g, Error &err) {
    const NodeDecl *voices = nullptr, *handlers = nullptr, *alloc = nullptr, *parser = nullptr, *post = nullptr;
    for (const NodeDecl &d : top.nodes) {
        if (d.rate) { err.at(d.at, "node rates on the nodes of a polyphonic graph are not supported (oversample inside the voice graph)"); return false; }
        if (d.type == "MidiParser" && !d.array) { parser = &d; continue; }
        /* the instance-level tail: ONE Tremolo between the voice sum and a Frame<2> output (examples/electric-piano/src/main.rs:58,88-96) */
        if (d.type == "Tremolo" && !d.array && !post) {
            if (d.ctor != "new" || !d.args.empty()) { err.at(d.at, "Tremolo is constructed with `Tremolo::new()`"); return false; }
            post = &d; continue;
        }
        if (d.type == "VoiceAllocator" && !d.array) { alloc = &d; continue; }
        if (d.type == "MidiVoiceHandler" && d.array) { handlers = &d; continue; }
        if (d.array && find_graph(all, d.type) && !voices) { voices = &d; continue; }
        err.at(d.at, d.array ? "node array `" + d.name + "`: `" + d.type + "` is not a graph of this text (a polyphonic graph holds ONE array of a voice graph, plus MidiParser / VoiceAllocator / [MidiVoiceHandler; N])"
                             : "node `" + d.name + "`: a polyphonic graph holds one voice array plus MidiParser / VoiceAllocator / [MidiVoiceHandler; N] and at most one Tremolo after the voice sum");
        return false;
    }
    if (!voices) { Tok t; err.at(t, "polyphonic graph without a voice array"); return false; }
    const Ast &voice = *find_graph(all, voices->type);
    if (is_polyphonic(voice)) { err.at(voices->at, "the voice graph `" + voice.name + "` is itself polyphonic: one level of node arrays is supported"); return false; }
    if (voices->ctor != "new" || !voices->args.empty()) { err.at(voices->at, "a graph is constructed with `" + voice.name + "::new()`"); return false; }
    if ((uint32_t)voices->array > 64u || (voices->array & (voices->array - 1)) != 0) { err.at(voices->at, "voice array size must be a power of two <= 64 (the engine folds an instance inside one 64-thread CTA)"); return false; }
    if (handlers && handlers->array != voices->array) { err.at(handlers->at, "`" + handlers->name + "` and `" + voices->name + "` must have the same array size"); return false; }
    {
        Ast flat_voice;                                   /* the voice graph may itself use nested graphs */
        Inliner inl{all, err};
        if (!inl.flatten(voice, flat_voice)) return false;
        Lowering l(flat_voice, g, err);
        if (!l.run()) return false;
    }
    g.vpi = (uint32_t)voices->array;
    g.midi = parser && alloc && handlers;
    if ((parser || alloc || handlers) && !g.midi) { Tok t; err.at(t, "the MIDI front needs all of MidiParser, VoiceAllocator and [MidiVoiceHandler; N]"); return false; }
    const std::string voice_name = g.name;
    g.name = top.name;
    std::map<std::string, const InputDecl *> outer_in;
    std::string outer_out, outer_midi;
    for (const InputDecl &d : top.inputs) {
        outer_in[d.name] = &d;
        if (d.kind == Kind::Event) outer_midi = d.name;
        else if (d.kind != Kind::Value) { err.at(d.at, "a polyphonic graph takes value inputs (broadcast to the voices) and the MIDI event input"); return false; }
    }
    std::vector<std::string> event_outs;                  /* `output note_on_out: event;`: accepted, nothing on the device consumes them */
    int out_lanes = 1;
    for (const OutputDecl &d : top.outputs) {
        if (d.kind == Kind::Event) { event_outs.push_back(d.name); continue; }
        if (d.kind != Kind::Stream || !outer_out.empty()) { err.at(d.at, "a polyphonic graph has one stream output (the voice sum)"); return false; }
        outer_out = d.name; out_lanes = d.lanes;
        if (d.lanes == 2 && !post) { err.at(d.at, "a `Frame<2>` output needs a node that produces one: `Tremolo` after the voice sum"); return false; }
        if (d.lanes == 1 && post) { err.at(d.at, "Type mismatch in connection: `Tremolo.output` is Frame<2> but `" + d.name + "` is f32 (declare `output " + d.name + ": stream: Frame<2>;`)"); return false; }
    }
    g.post_tremolo = post != nullptr;
    g.channels = (uint32_t)out_lanes;
    auto voice_value_id = [&](const std::string &name) -> int {
        for (size_t i = 0; i < g.inputs.size(); ++i) if (g.inputs[i].name == name) return (int)i;
        return -1;
    };
    auto is_node = [&](const Expr &e, const NodeDecl *nd, std::string &field) {
        if (!nd || e.k != Expr::Field || e.a->k != Expr::Ident || e.a->name != nd->name) return false;
        field = e.name; return true;
    };
    bool freq_wired = false, gate_wired = false, out_wired = false, post_in = false, post_out = false;
    for (const Conn &c : top.conns) {
        if (c.via) { err.at(c.at, "feedback edges are not supported in a polyphonic graph"); return false; }
        std::string sf, df;
        const Expr &s = *c.src, &d = *c.dst;
        /* voices.<out> -> out : the in-order voice sum (the engine's instance fold) */
        if (!post && is_node(s, voices, sf) && d.k == Expr::Ident && d.name == outer_out) {
            if (sf != g.output_name) { err.at(s.at, "`" + voice.name + "` has no stream output `" + sf + "`"); return false; }
            out_wired = true; continue;
        }
        /* voices.<out> -> tremolo.input (the voice sum, stream fan-in of an array source: emit_node.rs:463-466), tremolo.output -> out,
           value -> tremolo.rate / .depth (instance-level value inputs: they get ids after the voice's own) */
        if (post && is_node(d, post, df)) {
            if (df == "input") {
                if (!is_node(s, voices, sf) || sf != g.output_name) { err.at(s.at, "only the voice sum `" + voices->name + "." + g.output_name + "` can feed `" + post->name + ".input`"); return false; }
                post_in = true; continue;
            }
            const int k = df == "rate" ? 0 : df == "depth" ? 1 : -1;
            if (k < 0) { err.at(d.at, "node `" + post->name + "` has no endpoint `" + df + "` (Tremolo: input, rate, depth -> output)"); return false; }
            if (s.k == Expr::Literal) { g.post_pid[k] = plan::kNoValue; g.post_cst[k] = s.lit; continue; }
            if (s.k == Expr::Ident && outer_in.count(s.name) && outer_in[s.name]->kind == Kind::Value) {
                if (g.inputs.size() >= 28) { err.at(s.at, "more than 28 value inputs (OSCEN_B200_MAX_INPUTS)"); return false; }
                const InputDecl &od = *outer_in[s.name];
                plan::ValueInput spec = od.spec; spec.name = s.name;
                if (!od.has_def) spec.def = 0.0f;
                g.post_pid[k] = (uint32_t)g.inputs.size();
                g.inputs.push_back(spec);
                continue;
            }
            err.at(s.at, "only a value input of the outer graph or a literal can feed `" + post->name + "." + df + "`"); return false;
        }
        if (post && is_node(s, post, sf)) {
            if (sf != "output" || d.k != Expr::Ident || d.name != outer_out) { err.at(c.at, "`" + post->name + ".output` connects to the graph's stream output"); return false; }
            post_out = true; continue;
        }
        /* parser.note_on / note_off -> an event output of the graph: a host-side tap with no consumer in the engine */
        if (is_node(s, parser, sf) && d.k == Expr::Ident && std::find(event_outs.begin(), event_outs.end(), d.name) != event_outs.end()) {
            if (sf != "note_on" && sf != "note_off") { err.at(s.at, "MidiParser has no output `" + sf + "` (note_on, note_off)"); return false; }
            continue;
        }
        if (is_node(d, voices, df)) {
            /* handlers.frequency / handlers.gate -> voices.<input> */
            if (is_node(s, handlers, sf)) {
                if (sf == "frequency") {
                    const int id = voice_value_id(df);
                    if (id < 0) { err.at(d.at, "`" + voice.name + "` has no value input `" + df + "`"); return false; }
                    g.midi_freq_input = (uint32_t)id; freq_wired = true; continue;
                }
                if (sf == "gate") {
                    if (df != g.event_input) { err.at(d.at, "`" + voice.name + "` has no event input `" + df + "`"); return false; }
                    gate_wired = true; continue;
                }
                err.at(s.at, "MidiVoiceHandler has no output `" + sf + "` (frequency, gate)"); return false;
            }
            /* outer value input (or literal) -> voices.<value input>: broadcast */
            const int id = voice_value_id(df);
            if (id < 0) {
                err.at(d.at, df == g.event_input ? "Type mismatch in connection: source is Value but destination expects Event"
                                                 : "`" + voice.name + "` has no value input `" + df + "`");
                return false;
            }
            if (s.k == Expr::Literal) { g.inputs[id].def = s.lit; continue; }
            if (s.k == Expr::Ident && outer_in.count(s.name) && outer_in[s.name]->kind == Kind::Value) {
                const InputDecl &od = *outer_in[s.name];
                plan::ValueInput spec = od.spec;              /* the outer declaration rules: default, range, ramp */
                spec.name = g.inputs[id].name;
                if (!od.has_def) spec.def = 0.0f;
                g.inputs[id] = spec;
                g.aliases.emplace_back(s.name, (uint32_t)id);
                continue;
            }
            err.at(s.at, "only a value input of the outer graph, a literal or a MidiVoiceHandler output can feed `" + voices->name + "." + df + "`"); return false;
        }
        /* the MIDI chain itself: midi_in -> parser -> allocator -> handlers */
        const bool midi_edge =
            (s.k == Expr::Ident && s.name == outer_midi && is_node(d, parser, df) && df == "midi_in") ||
            (is_node(s, parser, sf) && is_node(d, alloc, df) && sf == df && (sf == "note_on" || sf == "note_off")) ||
            (is_node(s, alloc, sf) && sf == "voices" && is_node(d, handlers, df) && (df == "note_on" || df == "note_off"));
        if (midi_edge) continue;
        err.at(c.at, "this connection is not part of the polyphonic pattern (MIDI chain, broadcast `value -> voices.input`, `voices.output -> out`)");
        return false;
    }
    if (post) {
        if (!post_in || !post_out) { err.at(post->at, "`" + post->name + "` must sit between the voice sum and the output (`" + voices->name + "." + g.output_name + " -> " + post->name + ".input; " + post->name + ".output -> " + outer_out + ";`)"); return false; }
        out_wired = true;
    }
    if (!out_wired) { Tok t; err.at(t, "`" + voices->name + "." + g.output_name + "` is not connected to the output"); return false; }
    if (g.midi && !(freq_wired && gate_wired)) { Tok t; err.at(t, "the MIDI front is declared but `" + handlers->name + ".frequency` / `.gate` are not connected to the voices"); return false; }
    g.output_name = outer_out + (post ? " <- Tremolo (Frame<2>)" : "") + " <- sum of " + std::to_string(g.vpi) + " x " + voice_name + "." + g.output_name;
    return true;
}

/* text -> plan::Graph; on failure `error` holds "line:col: message".  With several `graph! { }` blocks the
   LAST one is the graph to run; earlier ones are the graphs it instantiates. */
inline bool compile(const char *text, plan::Graph &g, std::string &error, const char *top_name = nullptr) {
    Error err;
    std::vector<Tok> toks;
    if (!text) { error = "0:0: null graph text"; return false; }
    if (!lex(text, toks, err)) { error = err.text(); return false; }
    if (!expand_oversample_variants(toks, err)) { error = err.text(); return false; }
    std::vector<Ast> graphs;
    Parser p(toks, err);
    if (!p.parse(graphs)) { error = err.text(); return false; }
    if (graphs.empty()) { error = "1:1: no graph in the text"; return false; }
    { ArrayExpander ax{err, {}}; for (Ast &a : graphs) if (!ax.run(a)) { error = err.text(); return false; } }
    const Ast *chosen = &graphs.back();
    if (top_name && *top_name) {
        chosen = find_graph(graphs, top_name);
        if (!chosen) {
            std::string have;
            for (const Ast &a : graphs) have += (have.empty() ? "" : ", ") + (a.name.empty() ? std::string("<anonymous>") : a.name);
            error = "1:1: no graph named `" + std::string(top_name) + "` in the text (it holds: " + have + ")"; return false;
        }
    }
    const Ast &top = *chosen;
    if (is_polyphonic(top)) {
        if (!lower_polyphonic(top, graphs, g, err)) { error = err.text(); return false; }
        return true;
    }
    Ast flat;
    Inliner inl{graphs, err};
    if (!inl.flatten(top, flat)) { error = err.text(); return false; }
    Lowering l(flat, g, err);
    if (!l.run()) { error = err.text(); return false; }
    return true;
}

/* text form of a compiled graph: inputs, then the launch plan (for tests and tools) */
inline std::string describe(const plan::Graph &g, const plan::Plan &p) {
    std::string s = "graph " + (g.name.empty() ? std::string("<anonymous>") : g.name) + "\n";
    for (size_t i = 0; i < g.inputs.size(); ++i) {
        char b[160];
        snprintf(b, sizeof b, "input %zu: %s = %.9g", i, g.inputs[i].name.c_str(), (double)g.inputs[i].def);
        s += b;
        if (g.inputs[i].has_range) { snprintf(b, sizeof b, " [%.9g..%.9g%s]", (double)g.inputs[i].lo, (double)g.inputs[i].hi, g.inputs[i].log ? ", log" : ""); s += b; }
        if (g.inputs[i].ramp) s += " ramp " + std::to_string(g.inputs[i].ramp);
        s += "\n";
    }
    for (const auto &al : g.aliases) s += "alias: " + al.first + " -> input " + std::to_string(al.second) + "\n";
    if (g.vpi) s += "polyphony: " + std::to_string(g.vpi) + " voices per instance" + (g.midi ? ", MIDI front -> frequency = input " + std::to_string(g.midi_freq_input) + ", gate" : "") + "\n";
    if (!g.stream_input.empty()) s += "stream input: " + g.stream_input + "\n";
    if (!g.event_input.empty()) s += "event input: " + g.event_input + "\n";
    for (size_t k = 0; k < g.assets.size(); ++k) {
        s += "external " + std::to_string(k) + ": " + g.assets[k] + " (AudioAsset) ->";
        bool any = false;
        for (const plan::Node &nd : g.nodes) if (nd.asset == (int)k) { s += std::string(any ? ", " : " ") + nd.name + ".buf"; any = true; }
        s += any ? "\n" : " (not connected)\n";
    }
    s += "output: " + g.output_name + " <- " + g.nodes[g.output.node].name + "\n";
    s += "state words per voice: " + std::to_string(g.n_state) + "\n";
    for (size_t k = 0; k < p.launches.size(); ++k) {
        const plan::Launch &l = p.launches[k];
        s += std::to_string(k) + ": " + plan::type_name(l.type);
        s += " [";
        for (size_t j = 0; j < l.nodes.size(); ++j) {
            if (j) s += ", ";
            const plan::Node &nd = g.nodes[l.nodes[j]];
            s += nd.name;
            if (nd.type == plan::NodeType::Expr) s += std::string(" {") + plan::op_name(nd.op) + "}";
        }
        s += "]\n";
    }
    return s;
}

}  // namespace dsl
}  // namespace oscen
