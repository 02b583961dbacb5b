#!/bin/bash
# Evidence that the shipped library is sm_100a-native: which SASS mnemonics lib/libfmcuda.so contains.
# usage: bash profiles/sass_summary.sh > profiles/sass_summary_r2.txt   (no GPU needed)
LIB=finmath-lib-cuda-extensions_b200/lib/libfmcuda.so
echo "# cuobjdump of $LIB ($(date -u +%Y-%m-%d)); counts of SASS mnemonics over all kernels"
echo "# architectures in the fat binary:"; cuobjdump -lelf $LIB | sed 's/^/#   /'
cuobjdump -sass $LIB > /tmp/fmcuda.sass
for m in UBLKCP UTMALDG "SYNCS.PHASECHK" "SYNCS.ARRIVE" BRX ELECT MUFU.RCP FMNMX3 "LDS.128" "STS.128" "STG.E.128" "LDG.E.128" REDUX SHFL FCHK HMMA; do
  printf "%-16s %8d\n" "$m" "$(grep -c -- "$m" /tmp/fmcuda.sass)"
done
echo "# kernels:"; grep "Function :" /tmp/fmcuda.sass | sed 's/.*Function : /#   /' | c++filt | sort | uniq -c | sort -rn | head -60
