#!/bin/bash
O=gpurun_out/${1:-slots2}
mkdir -p $O
SEQALK_OVERLAP=0 python scripts/shard_span.py cfg4 8 14 2 2>&1 | sed "s/^/nooverlap /" | tee -a $O/span.txt
SEQALK_SLOTS=2 python scripts/shard_span.py cfg4 8 14 3 2>&1 | sed "s/^/slots=2 /" | tee -a $O/span.txt
SEQALK_SLOTS=3 SEQALK_DIR_BUDGET_GB=40 python scripts/shard_span.py cfg4 8 14 3 2>&1 | sed "s/^/slots=3 budget40 /" | tee -a $O/span.txt
SEQALK_SLOTS=4 SEQALK_DIR_BUDGET_GB=38 python scripts/shard_span.py cfg4 8 14 3 2>&1 | sed "s/^/slots=4 budget38 /" | tee -a $O/span.txt
SEQALK_SLOTS=2 python scripts/shard_span.py cfg4 8 10 3 2>&1 | sed "s/^/slots=2 10chunks /" | tee -a $O/span.txt
SEQALK_SLOTS=3 SEQALK_DIR_BUDGET_GB=40 python scripts/shard_span.py cfg4 8 10 3 2>&1 | sed "s/^/slots=3 budget40 10chunks /" | tee -a $O/span.txt
