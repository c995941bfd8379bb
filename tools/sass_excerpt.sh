#!/bin/bash
# Per-kernel count of the Blackwell-native SASS mnemonics in the in-tree library (B200_PROFILING.md, "What proves a Blackwell-native
# kernel"): UTCHMMA = tcgen05.mma, UTMALDG = tiled TMA load, UBLKCP = 1-D bulk copy, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit.
# Runs on the build container (no GPU needed).   usage: tools/sass_excerpt.sh > profiles/rNN_sass_excerpt.md
SO=${1:-tandem_b200/libtandem_b200.so}
echo "# SASS evidence: tcgen05 / TMEM / TMA mnemonics per kernel of \`$SO\`"
echo
echo "\`cuobjdump -sass $SO\` (sm_100a), mnemonic counts per kernel; kernels without any of them (the bandwidth / latency kernels) are listed at the end."
echo
echo "| kernel | UTCHMMA (tcgen05.mma) | UTMALDG (TMA tile) | UBLKCP (bulk copy) | LDTM (tcgen05.ld) | STTM (tcgen05.st) | UTCBAR (commit) | SYNCS (mbarrier) |"
echo "|---|---|---|---|---|---|---|---|"
cuobjdump -sass "$SO" | awk '
  /Function :/ { name=$3; sub(/^[ \t]+/, "", name); order[++n]=name }
  /UTCHMMA/ {a[name]++} /UTMALDG/ {b[name]++} /UBLKCP/ {c[name]++} /LDTM/ {d[name]++} /STTM/ {e[name]++} /UTCBAR/ {f[name]++} /SYNCS/ {g[name]++}
  END { for (i=1;i<=n;i++) { k=order[i]; if (a[k]+b[k]+c[k]+d[k]+e[k]+f[k] > 0) printf "%s\t%d\t%d\t%d\t%d\t%d\t%d\t%d\n", k, a[k], b[k], c[k], d[k], e[k], f[k], g[k];
                              else plain[++m]=k }
        for (i=1;i<=m;i++) print "PLAIN\t" plain[i] }' > /tmp/sass_counts.txt
grep -v "^PLAIN" /tmp/sass_counts.txt | while IFS=$'\t' read -r k a b c d e f g; do
  echo "| \`$(echo "$k" | c++filt | sed 's/|/\\|/g' | cut -c 1-150)\` | $a | $b | $c | $d | $e | $f | $g |"
done
echo
echo "Totals: $(grep -v '^PLAIN' /tmp/sass_counts.txt | awk -F'\t' '{a+=$2;b+=$3;c+=$4;d+=$5;e+=$6;f+=$7} END {printf "%d UTCHMMA, %d UTMALDG, %d UBLKCP, %d LDTM, %d STTM, %d UTCBAR", a,b,c,d,e,f}') in $(grep -vc '^PLAIN' /tmp/sass_counts.txt) kernels."
echo
echo "No \`HMMA\` (legacy mma.sync) anywhere: $(cuobjdump -sass "$SO" | grep -c ' HMMA') occurrences."
echo
echo "## Kernels without tensor-core / TMA instructions (bandwidth-, issue- or latency-bound by design)"
echo
grep "^PLAIN" /tmp/sass_counts.txt | cut -f2 | c++filt | sed 's/(.*//' | sort -u | sed 's/^/- `/; s/$/`/'
