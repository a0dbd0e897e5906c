for rpb in 8 4 2; do for pm in 0 64; do echo "== rpb=$rpb persist=$pm"; SQ_ROWS_PER_BIN=$rpb SQ_L2_PERSIST_MB=$pm ./tools/bench_quick.sh "cfg5_shard"; done; done
