for d in 16 17 21 23 0 1; do echo "PN_TC_DEBUG=$d"; PN_TC_DEBUG=$d python scripts/tc_probe.py 1000000 151552 16 2>&1 | grep "k=1:"; done
