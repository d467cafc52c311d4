set -x
O=gpurun_out
for b in 0 5 7 8 9 10 12 14 16 18; do
  echo "== bands $b" >> $O/r3i_lat.txt
  B200_CANNY_BANDS=$b timeout 300 python tools/pdl_probe.py >> $O/r3i_lat.txt 2>> $O/r3i.err
done
cat $O/r3i_lat.txt | cut -c1-420
