VBK_PROF=1 VBK_LOOKAHEAD=0 python profiles/fast_one.py dfl001 2>&1 | grep -i "profile" | tail -1
