# developer sweep: K3 column form (G,RPT)
for v in 1,32 2,16 4,8 4,16 8,8; do echo "== $v"; QNMFIT_K3G=$v python tools/k3_time.py 5 2>&1 | tail -1 | cut -c1-100; done
