#!/bin/bash
# usage: tools/build_rev.sh <git-rev> <name>   -> dispersal-simulation_b200/ab/<name>.so  (library of another revision, for tools/ab.sh A/B runs)
set -e
rev=$1; name=$2
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
git -C "$root" archive "$rev" dispersal-simulation_b200 include | tar -x -C "$tmp"
make -s -C "$tmp/dispersal-simulation_b200" libdsim_b200.so > /dev/null 2>&1
mkdir -p "$root/dispersal-simulation_b200/ab"
cp "$tmp/dispersal-simulation_b200/libdsim_b200.so" "$root/dispersal-simulation_b200/ab/$name.so"
rm -rf "$tmp"
echo "built $name.so from $rev"
