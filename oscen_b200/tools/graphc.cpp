// graphc — the graph compiler as a command-line tool (host only: needs neither CUDA nor a GPU).
//
//   graphc [--plan | --cuda | --check] [--graph NAME] file.graph [more.graph ...]
//   graphc --nodes                      the node library: constructors, endpoints, defaults
//
// Concatenates the files (a voice graph followed by the polyphonic graph that instantiates it, say), runs the
// `graph!` front end (oscen_b200/csrc/dsl.hpp) and prints the lowered plan (--plan, default), the CUDA source of
// the generated fused kernel (--cuda), or nothing (--check).  Diagnostics go to stderr as `line:col: message`;
// exit code 1 on a compile error, 2 on a usage / IO error.  The same code runs inside liboscen_b200.so behind
// oscen_b200_graph_compile / _describe / _codegen; this tool exists for offline inspection and for running the
// front end under sanitizers (tests/test_dsl_cpu.py builds it with -fsanitize=address,undefined).
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>

#include "../csrc/codegen.hpp"
#include "../csrc/dsl.hpp"

int main(int argc, char **argv) {
    std::string mode = "--plan", text, top;
    int n_files = 0;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--nodes") { std::fputs(oscen::dsl::node_library_text().c_str(), stdout); return 0; }     /* the node library the front end knows */
        if (a == "--plan" || a == "--cuda" || a == "--check") { mode = a; continue; }
        if (a == "--graph" && i + 1 < argc) { top = argv[++i]; continue; }     /* compile this graph instead of the last one */
        if (a == "-") { std::stringstream ss; ss << std::cin.rdbuf(); text += ss.str(); ++n_files; continue; }
        std::ifstream f(a, std::ios::binary);
        if (!f) { std::fprintf(stderr, "graphc: cannot read %s\n", a.c_str()); return 2; }
        std::stringstream ss; ss << f.rdbuf();
        text += ss.str(); text += "\n"; ++n_files;
    }
    if (!n_files) { std::fprintf(stderr, "usage: graphc [--plan | --cuda | --check] [--graph NAME] file.graph [more.graph ...]   (`-` = stdin)\n"); return 2; }
    oscen::plan::Graph g;
    oscen::plan::Plan p;
    std::string err;
    if (!oscen::dsl::compile(text.c_str(), g, err, top.c_str())) { std::fprintf(stderr, "%s\n", err.c_str()); return 1; }
    if (!oscen::plan::compile(g, p, err)) { std::fprintf(stderr, "0:0: %s\n", err.c_str()); return 1; }
    if (mode == "--plan") std::fputs(oscen::dsl::describe(g, p).c_str(), stdout);
    else if (mode == "--cuda") std::fputs(oscen::codegen::fused_kernel_source(g, p).c_str(), stdout);
    return 0;
}
