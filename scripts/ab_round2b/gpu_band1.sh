#!/bin/bash
# first GPU call of the banded traceback + host hashing: parity tests, then same-box A/B on cfg5 / cfg2
O=gpurun_out/band1
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "band or BAND or HOST_HASH or cfg5_full_size_golden or golden_big13" 2>&1 | tail -15 | tee $O/pytest.txt
export SEQALK_OVERLAP=0
timeout 300 python scripts/ab_env.py cfg5 3 "SEQALK_TB_BAND=0,SEQALK_HOST_HASH=0" "" "SEQALK_TB_BAND_HALFWIDTH=1" "SEQALK_TB_BAND_HALFWIDTH=2" \
   "SEQALK_TB_BAND_LEVELS=31" "SEQALK_TB_BAND_LEVELS=127" "SEQALK_TB_BAND_LEVELS=255,SEQALK_TB_BAND_HALFWIDTH=8" 2>&1 | tee $O/ab_cfg5.txt
timeout 300 python scripts/ab_env.py cfg2 3 "SEQALK_TB_BAND=0,SEQALK_HOST_HASH=0" "" "SEQALK_TB_BAND_HALFWIDTH=1" "SEQALK_TB_BAND_HALFWIDTH=4" \
   "SEQALK_TB_BAND_LEVELS=31" "SEQALK_TB_BAND_LEVELS=127" "SEQALK_HOST_HASH=1" 2>&1 | tee $O/ab_cfg2.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:band -c 120 --csv --log-file $O/band_launches.csv \
   python scripts/ab_env.py cfg5 1 "" > $O/ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/band1/band_launches.csv")) if len(r) > 5 and r[0].isdigit()]
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    k = "produce" if "produce" in r[4] else "walk"
    v = float(r[-1].replace(",", "")); u = r[-2]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    tot[k][0] += 1; tot[k][1] += ms
for k, (n, ms) in tot.items():
    print(f"ncu (cold, serialised): {k}: {n} launches, {ms:.3f} ms in total")
PY
