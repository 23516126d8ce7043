python tools/profile_train.py 64 2>&1 | tail -24
