for v in X1 X1DVBK_X2; do
cp linear-programming-vanderbei_b200/libvbkkt.so /tmp/orig.so
cp build/libvbkkt_$v.so linear-programming-vanderbei_b200/libvbkkt.so
echo variant $v
VBK_PROF=1 VBK_LOOKAHEAD=0 python profiles/fast_one.py dfl001 2>&1 | grep -i "profile" | tail -1
cp /tmp/orig.so linear-programming-vanderbei_b200/libvbkkt.so
done
