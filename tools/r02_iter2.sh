# iteration + grouped profile in one call: $1 = tag
tag=$1
bash tools/r02_iter.sh $tag ${2:-28416}
bash tools/r02_prof_g.sh ${tag}_hammer 5
