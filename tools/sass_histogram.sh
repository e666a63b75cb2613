#!/bin/bash
# SASS evidence of which kernels are Blackwell-native: per kernel of lib/libdquartic_b200.so the count of tcgen05 (UTC*MMA),
# TMEM (LDTM / STTM), TMA / bulk-copy (UTMALDG / UBLKCP), legacy tensor (HMMA), MUFU and packed-FP32 (FFMA2 / FADD2 / FMUL2)
# instructions.  Runs here (no GPU): tools/sass_histogram.sh > profiles/r2_sass_histogram.txt
cd "$(dirname "$0")/.."
lib=diffusion-deconvolution-dia-msms-data_b200/lib/libdquartic_b200.so
cuobjdump -sass "$lib" | awk '
  /Function :/ { fn=$3 }
  /UTC[A-Z]*MMA/ { utc[fn]++ } /LDTM/ { ldtm[fn]++ } /STTM/ { sttm[fn]++ } /UTMALDG|UBLKCP/ { tma[fn]++ }
  /HMMA/ { hmma[fn]++ } /MUFU/ { mufu[fn]++ } /FFMA2|FADD2|FMUL2/ { pk[fn]++ } /UTCBAR/ { bar[fn]++ }
  { if (fn != "") n[fn]++ }
  END { printf "%-8s %-6s %-6s %-8s %-6s %-6s %-8s %-8s %s\n", "UTC*MMA", "LDTM", "STTM", "TMA/BLK", "HMMA", "MUFU", "F*2", "instr", "kernel";
        for (f in n) if (utc[f] + ldtm[f] + tma[f] + hmma[f] > 0)
          printf "%-8d %-6d %-6d %-8d %-6d %-6d %-8d %-8d %s\n", utc[f], ldtm[f], sttm[f], tma[f], hmma[f], mufu[f], pk[f], n[f], f }' |
  (read h; echo "$h"; sort -k9 | c++filt | cut -c1-220)
