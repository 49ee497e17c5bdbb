// check_cells.cpp — runs the C++ constraint re-checker (include/paillier_chip_host.hpp: check_constraints) over a cell stream
// read from a text file, so that the host logic can be tested without a GPU: tests/test_host_cpp.py writes the file from the
// chip restatement's cells (oracle) and expects "satisfied", then flips cells and expects "violated".
// File format (hex without 0x, one record per line):
//   H <limb_bits> <lookup_bits>
//   C <cell>                              one line per advice cell, in assignment order
//   A <first_cell> <n_limbs> <value>      an assign_integer
//   N <first_cell> <n_limbs> <n>          a square+refresh of n
//   M <first_cell> <L> <a> <b> <n2>       a mul_mod group
// Exit code 0 = satisfied, 3 = violated (the violations are printed), 1 = bad input.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include "paillier_chip_host.hpp"

using namespace paillier_halo2;

static BigUint from_hex(const std::string& h) {
    BigUint r;
    for (char ch : h) {
        int v = ch >= '0' && ch <= '9' ? ch - '0' : ch >= 'a' && ch <= 'f' ? ch - 'a' + 10 : ch >= 'A' && ch <= 'F' ? ch - 'A' + 10 : -1;
        if (v < 0) throw std::runtime_error("bad hex digit");
        r = (r << 4) + BigUint((uint64_t)v);
    }
    return r;
}
static AssignedBigUint limbs_of(const BigUint& v, size_t nl, unsigned limb_bits) {
    AssignedBigUint a; a.int_value = v; a.max_limb_bits = limb_bits;
    for (size_t l = 0; l < nl; l++) a.limb_values.push_back((v >> (l * limb_bits)).low_bits(limb_bits));
    return a;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: check_cells FILE\n"); return 1; }
    std::ifstream in(argv[1]);
    if (!in) { fprintf(stderr, "cannot open %s\n", argv[1]); return 1; }
    Context ctx;
    unsigned limb_bits = 0, lookup_bits = 0;
    std::string line;
    try {
        while (std::getline(in, line)) {
            std::istringstream ss(line);
            std::string tag; ss >> tag;
            if (tag == "H") ss >> limb_bits >> lookup_bits;
            else if (tag == "C") { std::string h; ss >> h; BigUint v = from_hex(h); Fr f{0, 0, 0, 0}; v.to_words(f.data(), 4); ctx.cells.push_back(f); }
            else if (tag == "A") { size_t first, nl; std::string h; ss >> first >> nl >> h; ctx.assigns.push_back(AssignRecord{first, nl, from_hex(h)}); }
            else if (tag == "N") { size_t first, nl; std::string h; ss >> first >> nl >> h; ctx.n2s.push_back(N2Record{first, limbs_of(from_hex(h), nl, limb_bits)}); }
            else if (tag == "M") {
                size_t first, L; std::string a, b, n2; ss >> first >> L >> a >> b >> n2;
                ctx.mul_mods.push_back(MulModGroup{first, limbs_of(from_hex(a), L, limb_bits), limbs_of(from_hex(b), L, limb_bits), limbs_of(from_hex(n2), L, limb_bits)});
            }
        }
        std::vector<std::string> bad = check_constraints(ctx, limb_bits, lookup_bits);
        printf("%zu cells, %zu assigns, %zu n2, %zu mul_mods: %s\n", ctx.cells.size(), ctx.assigns.size(), ctx.n2s.size(), ctx.mul_mods.size(),
               bad.empty() ? "satisfied" : "violated");
        for (size_t i = 0; i < bad.size() && i < 8; i++) printf("  %s\n", bad[i].c_str());
        return bad.empty() ? 0 : 3;
    } catch (const std::exception& e) { fprintf(stderr, "error: %s\n", e.what()); return 1; }
}
