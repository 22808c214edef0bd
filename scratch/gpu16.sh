VBK_PROF=1 python profiles/fast_one.py dfl001 2>&1 | grep "panel profile" | tail -2
