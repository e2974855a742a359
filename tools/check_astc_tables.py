"""Developer tool: cross-check the spec-derived ASTC tables of cuttlefish_b200/csrc/astc_tables.hpp
against the tables in the reference's astcenc sources (needs /root/reference)."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/lib/astc-encoder/Source/"
dump = r'''
#include "../cuttlefish_b200/csrc/astc_tables.hpp"
#include <cstdio>
using namespace cfx::astc;
int main() {
    for (int l = 0; l < kColorLevels; ++l) { printf("C %d", kColorQuant[l].n); for (int e = 0; e < kColorQuant[l].n; ++e) printf(" %d", unquant_color(e, kColorQuant[l])); printf("\n"); }
    for (int l = 0; l < kWeightLevels; ++l) { printf("W %d", kWeightQuant[l].n); for (int e = 0; e < kWeightQuant[l].n; ++e) printf(" %d", unquant_weight(e, kWeightQuant[l])); printf("\n"); }
    Built b = build_tables(6, 6);
    printf("T"); for (int i = 0; i < 243; ++i) printf(" %d", b.blob[b.tab.off_trit_enc + i]); printf("\n");
    printf("Q"); for (int i = 0; i < 125; ++i) printf(" %d", b.blob[b.tab.off_quint_enc + i]); printf("\n");
    printf("INFO grids %u modes1 %u modes2 %u part2 %u part3 %u blob %u\n", b.tab.n_grids, b.tab.n_modes1, b.tab.n_modes2, b.tab.n_part2, b.tab.n_part3, b.tab.blob_bytes);
    for (int s = 0; s < 1024; ++s) { printf("P2 %d", s); for (int i = 0; i < 36; ++i) printf(" %d", select_partition(s, i % 6, i / 6, 0, 2, false)); printf("\n"); }
    for (int s = 0; s < 1024; ++s) { printf("P3 %d", s); for (int i = 0; i < 36; ++i) printf(" %d", select_partition(s, i % 6, i / 6, 0, 3, false)); printf("\n"); }
    for (int s = 0; s < 1024; s += 7) { printf("S2 %d", s); for (int i = 0; i < 16; ++i) printf(" %d", select_partition(s, i % 4, i / 4, 0, 2, true)); printf("\n"); }
    return 0;
}
'''
cpp = os.path.join(HERE, "_build", "dump_astc.cpp")
open(cpp, "w").write(dump)
exe = os.path.join(HERE, "_build", "dump_astc")
subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", HERE, "-o", exe, cpp])
lines = subprocess.run([exe], capture_output=True, text=True).stdout.splitlines()
ours = {"C": {}, "W": {}}
for ln in lines:
    f = ln.split()
    if f[0] in ("C", "W"):
        ours[f[0]][int(f[1])] = [int(x) for x in f[2:]]
    elif f[0] == "T":
        trits = [int(x) for x in f[1:]]
    elif f[0] == "Q":
        quints = [int(x) for x in f[1:]]
    elif f[0] == "INFO":
        print(ln)

q = open(SRC + "astcenc_quantization.cpp").read()
bad = 0
for n, vals in ours["C"].items():
    m = re.search(r"color_scrambled_pquant_to_uquant_q%d\[%d\]\s*\{(.*?)\};" % (n, n), q, re.S)
    ref = [int(x) for x in re.findall(r"\d+", m.group(1))]
    if ref != vals:
        bad += 1
        print("colour level %d differs\n ours %s\n ref  %s" % (n, vals, ref))
w = open(SRC + "astcenc_weight_quant_xfer_tables.cpp").read()
for n, vals in ours["W"].items():
    m = re.search(r"//\s*QUANT_?%d,.*?\{\s*\{(.*?)\},\s*\{(.*?)\},\s*\{(.*?)\}," % n, w, re.S)
    sorted_vals = [int(x) for x in re.findall(r"\d+", m.group(1))]
    scramble = [int(x) for x in re.findall(r"\d+", m.group(2))]
    # rank k -> encoded integer scramble[k]; so encoded e -> value sorted_vals[rank of e]
    ref = [0] * n
    for k, e in enumerate(scramble):
        ref[e] = sorted_vals[k]
    if ref != vals:
        bad += 1
        print("weight level %d differs\n ours %s\n ref  %s" % (n, vals, ref))
s = open(SRC + "astcenc_integer_sequence.cpp").read()
m = re.search(r"integer_of_trits\[3\]\[3\]\[3\]\[3\]\[3\]\s*\{(.*?)\};", s, re.S)
ref = [int(x) for x in re.findall(r"\b\d+\b", m.group(1))]
# reference index order [t4][t3][t2][t1][t0] == ours t0 + 3 t1 + ...
if ref != trits:
    bad += 1
    print("trit table differs", sum(a != b for a, b in zip(ref, trits)))
m = re.search(r"integer_of_quints\[5\]\[5\]\[5\]\s*\{(.*?)\};", s, re.S)
ref = [int(x) for x in re.findall(r"\b\d+\b", m.group(1))]
if ref != quints:
    bad += 1
    print("quint table differs", sum(a != b for a, b in zip(ref, quints)))
print("table check:", "OK" if bad == 0 else "%d MISMATCHES" % bad)
open(os.path.join(HERE, "_build", "astc_parts.txt"), "w").write("\n".join(l for l in lines if l[0] in "PS"))
sys.exit(1 if bad else 0)
