cd $GRAFT_REPO_ROOT
gcc -std=c99 -I include examples/shard_host.c -L pydem_b200 -lpydem_b200 -lm -o /tmp/shard_host || exit 1
rm -f /tmp/nccl_id_c
for r in 0 1; do LD_LIBRARY_PATH=pydem_b200 timeout 120 /tmp/shard_host $r 2 /tmp/nccl_id_c & done; wait
