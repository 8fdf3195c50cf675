tag=r2c
out=gpurun_out
mkdir -p $out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file $out/${tag}_step_launches.csv python tools/profile_step.py > $out/${tag}_profile_step.log 2>&1
python tools/summarize_launches.py $out/${tag}_step_launches.csv ${tag}_step > $out/${tag}_summary_head.txt 2>&1
gzip -f -k $out/${tag}_step_launches.csv
ls -la $out | grep ${tag}_
