#!/bin/bash
# Per-kernel evidence that the hot kernels are Blackwell-native: the TMA / mbarrier / peer-store
# mnemonics in the SASS of libmgb.so (cuobjdump needs no GPU).  Writes profiles/sass_tile_kernels.txt
cd "$(dirname "$0")/.."
OUT=profiles/sass_tile_kernels.txt
SO=multigrid_parallel_b200/libmgb.so
{
echo "# SASS evidence, $(git rev-parse --short HEAD 2>/dev/null), $(date -u +%F): cuobjdump -sass $SO (sm_100a)"
echo "# per kernel: count of UTMALDG (cp.async.bulk.tensor load), UTMAPF (bulk tensor L2 prefetch), SYNCS (mbarrier),"
echo "# ST.E.*.STRONG.SYS / LD.E.*.STRONG.SYS (release / acquire at system scope: the fused halo flags), MEMBAR.*.SYS,"
echo "# DADD/DMUL (explicitly rounded fp64), DFMA (only inside division / reciprocal expansions and the verified fast quotient),"
echo "# HMMA/UTC*MMA (tensor cores: none on purpose)"
cuobjdump -sass "$SO" | awk '
/Function :/ { if (name != "") print_row(); name=$3; delete c; next }
{ if ($0 ~ /UTMALDG/) c["UTMALDG"]++; if ($0 ~ /UTMAPF/) c["UTMAPF"]++; if ($0 ~ /SYNCS/) c["SYNCS"]++;
  if ($0 ~ /STG?\.E.*STRONG\.SYS/) c["ST.SYS"]++; if ($0 ~ /LDG?\.E.*STRONG\.SYS/) c["LD.SYS"]++;
  if ($0 ~ /MEMBAR.*SYS/) c["MEMBAR.SYS"]++; if ($0 ~ / DADD /) c["DADD"]++; if ($0 ~ / DMUL /) c["DMUL"]++;
  if ($0 ~ / DFMA /) c["DFMA"]++; if ($0 ~ /HMMA|UTC.MMA/) c["MMA"]++; if ($0 ~ /SHFL/) c["SHFL"]++; }
function print_row() { printf "%-90s UTMALDG %2d UTMAPF %2d SYNCS %3d ST.SYS %2d LD.SYS %2d MEMBAR.SYS %2d DADD %4d DMUL %4d DFMA %3d SHFL %3d MMA %d\n", name, c["UTMALDG"], c["UTMAPF"], c["SYNCS"], c["ST.SYS"], c["LD.SYS"], c["MEMBAR.SYS"], c["DADD"], c["DMUL"], c["DFMA"], c["SHFL"], c["MMA"] }
END { print_row() }' | c++filt | sed 's/mgb::(anonymous namespace):://; s/mgb:://' | sort
echo
echo "# one UTMALDG / UTMAPF / SYNCS line of the dominant kernel (k_tile_sweep<1,6,43>):"
cuobjdump -sass -fun '_ZN3mgb12k_tile_sweepILi1ELi6ELi43EEEvNS_5TilePE14CUtensorMap_stS2_' "$SO" 2>/dev/null | grep -E "UTMALDG|UTMAPF|SYNCS|STRONG.SYS" | head -12
} > "$OUT"
wc -l "$OUT"
