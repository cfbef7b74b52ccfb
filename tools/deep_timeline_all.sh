for k in 6 7 8 9 10 11 12 13 14 15; do DUNET_DBG_LAUNCH=$k python tools/deep_timeline.py 2>&1 | tail -7; done
