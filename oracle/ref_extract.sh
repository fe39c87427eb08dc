#!/bin/bash
# ORACLE -- TEST INFRASTRUCTURE ONLY.  Prints lines A..B of a reference source file after checking that line A starts with
# the expected text (so that a different revision of the reference fails the build instead of compiling the wrong lines).
#   ref_extract.sh <file> <A> <B> <expected prefix of line A>
set -e
file=$1; a=$2; b=$3; want=$4
got=$(sed -n "${a}p" "$file")
case "$got" in
  "$want"*) ;;
  *) echo "ref_extract: $file:$a is '$got', expected it to start with '$want'" >&2; exit 1 ;;
esac
sed -n "${a},${b}p" "$file"
