#!/bin/bash
# Print registers / spills / smem of every kernel of ours from the ptxas logs.
cd "$(dirname "$0")/build"
for f in *.ptxas.log; do
  awk -v F="$f" '/Compiling entry function/ {name=$0; sub(/.*Compiling entry function ./,"",name); sub(/. for .*/,"",name)}
       /bytes stack frame/ {spill=$0; sub(/.*: */,"",spill)}
       /Used [0-9]+ registers/ {u=$0; sub(/.*Used /,"",u); if (name !~ /cub/) printf "%-18s %-90.90s | %s | %s\n", F, name, u, spill}' "$f"
done
