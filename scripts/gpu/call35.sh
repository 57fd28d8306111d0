for n in 1 2 4; do TGCN_T3_NI=$n python scripts/dbg_fwd_acc.py 2>&1 | grep "NI="; done
python scripts/dbg_fwd_acc.py 2>&1 | grep "NI="
TGCN_TC_V3=0 python scripts/dbg_fwd_acc.py 2>&1 | grep "NI=" | sed 's/^/v2: /'
