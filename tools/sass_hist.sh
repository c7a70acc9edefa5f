#!/bin/bash
# Per-kernel SASS opcode evidence for the shipped library (build container, no GPU): which kernels carry tcgen05 (UTC*MMA), TMEM
# loads (LDTM), TMA / bulk copies (UTMALDG, UBLKCP), cluster barriers (UCGABAR), warp shuffles, and whether anything spills (LDL/STL).
#   bash tools/sass_hist.sh > profiles/r2_sass_histogram.txt
cd "$(dirname "$0")/.."
LIB=rq_vae_recommender_b200/librqb200.so
echo "# cuobjdump -sass $LIB (built by __graft_entry__.build(): nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3), $(date -u +%F)"
cuobjdump -sass $LIB | awk '
  /Function :/ { name=$3; next }
  /^[ \t]*\/\*[0-9a-f]+\*\// {
    op=$2; if (op ~ /^@/) op=$3; sub(/;$/,"",op); split(op, parts, "."); base=parts[1];
    n[name]++;
    if (op ~ /^UTC.*MMA/) k[name,"UTC*MMA"]++;
    if (base=="LDTM") k[name,"LDTM"]++;
    if (base=="UTMALDG") k[name,"UTMALDG"]++;
    if (base=="UBLKCP") k[name,"UBLKCP"]++;
    if (base=="UTCBAR") k[name,"UTCBAR"]++;
    if (base ~ /^UCGABAR/) k[name,"UCGABAR"]++;
    if (base=="SYNCS") k[name,"SYNCS(mbarrier)"]++;
    if (base=="SHFL") k[name,"SHFL"]++;
    if (base=="VOTE") k[name,"VOTE"]++;
    if (base=="FFMA" || base=="FFMA2") k[name,"FFMA"]++;
    if (base=="LDL" || base=="STL") k[name,"LDL/STL"]++;
    if (base=="HMMA") k[name,"HMMA(legacy)"]++;
  }
  END {
    split("UTC*MMA LDTM UTMALDG UBLKCP UTCBAR UCGABAR SYNCS(mbarrier) SHFL VOTE FFMA LDL/STL HMMA(legacy)", cols, " ");
    printf "%-78s %7s", "kernel", "instrs"; for (c=1;c<=12;c++) printf " %9s", cols[c]; printf "\n";
    for (f in n) { printf "%-78s %7d", substr(f,1,78), n[f]; for (c=1;c<=12;c++) printf " %9d", k[f,cols[c]]+0; printf "\n" }
  }' | (read h1; echo "$h1"; sort)
