#!/bin/bash
# Static check of a build: per kernel code size, spills and how weights are fetched.  Per-thread register-indexed constant
# loads (LDC c[0x0][Rn+imm]) go through the indexed constant cache (~0.5 requests/clk/SM on B200) and must not appear in
# hot loops; uniform-register-indexed (LDCU c[0x0][URn+imm]) and immediate forms are fine.
# Usage: tools/sass_check.sh [lib.so] [kernel-regex]
LIB=${1:-gtcrn_micro_b200/libgtcrn_b200.so}; RE=${2:-.}
cuobjdump -sass "$LIB" 2>/dev/null | awk -v re="$RE" '/Function :/{name=$3} /c\[0x0\]\[R[0-9]/{a[name]++} /c\[0x0\]\[UR[0-9]/{b[name]++} /FFMA2/{f2[name]++} / FFMA /{f1[name]++} /\/\*[0-9a-f]+\*\/ +[A-Z@]/{cnt[name]++} END{for(n in cnt) if (n ~ re) printf "%6.1f KB  FFMA2 %5d FFMA %5d  LDC[R] %4d  LDCU[UR] %4d  %s\n", cnt[n]*16/1024, f2[n], f1[n], a[n], b[n], substr(n,1,70)}' | sort -k12
