# tuning builds of the wavefront TU: tools/build_variants.sh name "-DWF_X=.. -DWF_Y=.." [name2 "..."] -> digital-earth_b200/libde_<name>.so
set -e
cd "$(dirname "$0")/.."
while [ $# -ge 2 ]; do
  DE_LIB_SUFFIX=_$1 DE_WF_DEFS="$2" python digital-earth_b200/build.py --force > /dev/null
  echo "built libde_$1.so  [$2]"
  shift 2
done
python digital-earth_b200/build.py --force > /dev/null
