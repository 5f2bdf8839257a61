#!/bin/bash
# profiles/sass_summary.txt: per kernel of libflownet.so, how many tcgen05 / TMEM / TMA instructions its SASS holds
# (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = cp.async.bulk.tensor load / store,
#  UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier).  usage: tools/sass_summary.sh > profiles/sass_summary.txt
SO=${1:-steady-state-flow-with-neural-nets_b200/libflownet.so}
echo "# cuobjdump -sass $SO (sm_100a)"
echo "# kernel UTCHMMA LDTM STTM UTMALDG UTMASTG UBLKCP UTCBAR SYNCS"
cuobjdump -sass "$SO" | awk '
/Function :/ { if (name != "") print name, mma, ldtm, sttm, ldg, stg, blk, bar, syn; name=$3; mma=ldtm=sttm=ldg=stg=blk=bar=syn=0 }
/UTCHMMA/ {mma++} /LDTM/ {ldtm++} /STTM/ {sttm++} /UTMALDG/ {ldg++} /UTMASTG/ {stg++} /UBLKCP/ {blk++} /UTCBAR/ {bar++} /SYNCS/ {syn++}
END { print name, mma, ldtm, sttm, ldg, stg, blk, bar, syn }' | while read n rest; do echo "$(echo $n | c++filt | tr -d ' ' | cut -c1-120) $rest"; done | sort
