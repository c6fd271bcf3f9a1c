#!/bin/bash
# SHA-256 of the device code (SASS) of the built library, with the per-translation-unit hashes in the names of
# anonymous-namespace symbols masked.  Two builds with the same digest run the same instructions: a refactoring
# of host code or of qualifiers / comments in the headers can be committed without a GPU run when the digest
# stands (used for the host-callable geometry functions of round 2).   tools/sass_digest.sh [library]
LIB=${1:-$(dirname "$0")/../chess_b200/csrc/libchess_b200.so}
cuobjdump -sass "$LIB" \
  | sed -E 's/_GLOBAL__N__[0-9a-f]+_[0-9]+_[A-Za-z0-9_]+_cu_[0-9a-f]+/_ANON_/g; s/_INTERNAL_[0-9a-f]+_[0-9]+_[A-Za-z0-9_]+_cu_[0-9a-f]+/_INT_/g' \
  | grep -v "^\s*$" | sha256sum | cut -d' ' -f1
