for rpb in 8 2 1; do echo "== rows_per_bin $rpb"; RPB=$rpb python tools/time_subcfg.py 2>&1 >/dev/null | grep -E "cfg"; done
