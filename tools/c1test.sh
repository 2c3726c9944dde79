for v in tc cuda; do
SIR_CONV1_KERNEL=$v python tools/c1prof.py > /dev/null 2>&1 && SIR_CONV1_KERNEL=$v timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:conv1 -s 2 -c 2 --csv python tools/c1prof.py 2>/dev/null | grep -v "^==" | cut -d, -f5,11- | tail -9
done
