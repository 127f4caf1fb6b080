#!/bin/bash
# profiles/sass_excerpt.txt: per kernel of the shipped library, how many of the instructions the design rests on are in the SASS
#   STG.E.ENL2.256 (256-bit witness stores; in k_eval: value-slot stores) · LDG.E.ENL2.256 (k_eval value-slot loads) · SHFL (Keccak theta/rho-pi/chi lane exchanges, Poseidon mixes, prefix-sum scan)
#   UBLKCP (TMA bulk copy global->shared) · SYNCS (mbarrier arrive/expect_tx/try_wait) · UCGABAR (cluster barrier) · IMAD.WIDE (Fr limb products)
SO=proof-of-burn_b200/pob_b200/libpob_b200.so
{
echo "# cuobjdump -sass $SO  (sm_100a; $(date -u +%Y-%m-%dT%H:%MZ); $(python -c "import sys; sys.path.insert(0,'proof-of-burn_b200'); import pob_b200; print(pob_b200.lib().pob_version().decode())"))"
echo "# count  kernel  mnemonic"
cuobjdump -sass $SO | awk '/Function :/{f=$3} !/Function/{for(i=1;i<=NF;i++) if($i ~ /^(STG\.E\.ENL2\.256|LDG\.E\.ENL2\.256|SHFL\.[A-Z]+|UBLKCP\.S\.G|SYNCS\.[A-Z0-9.]+|UCGABAR_[A-Z]+|IMAD\.WIDE\.U32|UTMA[A-Z.]*|LDG\.E\.[0-9A-Z.]*CONSTANT)/){c[f" "$i]++}} END{for(k in c) print c[k], k}' | sort -k2,2 -k1,1nr | c++filt | sed -e 's/(anonymous namespace):://g' -e 's/void //' 
echo
echo "# first occurrences (address, encoding stripped)"
cuobjdump -sass $SO | grep -E "Function :|UBLKCP|SYNCS.ARRIVE.TRANS64|UCGABAR_ARV|STG.E.ENL2.256" | awk '/Function/{f=$0; n=0; next} {n++; if(n<=3){ if(f!=""){print f; f=""} sub(/\/\*[0-9a-f]+\*\/ *$/,""); print}}' | c++filt | sed -e 's/(anonymous namespace):://g' | cut -c1-170
} > profiles/sass_excerpt.txt
wc -l profiles/sass_excerpt.txt
