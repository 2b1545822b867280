#!/usr/bin/env python3
"""oracle/make_canonical.py -- TEST INFRASTRUCTURE.

Writes a patched COPY of the reference's ORBextractor.{h,cc} into a scratch
directory OUTSIDE the repository (never into the repo: reference sources are
not copied into history) with exactly the three "canonical tie-break" edits of
SURVEY.md section 8c:

  1. ExtractorNode gets a `long seq` member (include/ORBextractor.h:32-43);
  2. a per-call creation counter stamps every node right before it enters the
     list (src/ORBextractor.cc:561 push_back, :623..:720 the 8 push_front sites);
  3. the pointer-ordered sort (src/ORBextractor.cc:684) compares (size, seq).

Everything else stays byte-identical.  The reference sorts
pair<int, ExtractorNode*>, i.e. breaks ties in node size by heap address, which
makes its own output heap-dependent; the canonical variant is the definition of
"the reference's result" that all parity tests use.

usage: make_canonical.py <reference_root> <out_dir>
"""
import os
import re
import sys


def main():
    ref, out = sys.argv[1], sys.argv[2]
    os.makedirs(out, exist_ok=True)
    h = open(os.path.join(ref, "include", "ORBextractor.h")).read()
    cc = open(os.path.join(ref, "src", "ORBextractor.cc")).read()

    h2, n = re.subn(r"(\n\s*bool bNoMore;)", r"\1\n    long seq;", h, count=1)
    assert n == 1, "ExtractorNode::bNoMore not found"

    cc2, n = re.subn(r"(\n\s*list<ExtractorNode> lNodes;)", r"\n    long orbxSeqCounter = 0;\1", cc, count=1)
    assert n == 1, "lNodes declaration not found"
    cc2, n = re.subn(r"lNodes\.push_back\(ni\);", "ni.seq = orbxSeqCounter++; lNodes.push_back(ni);", cc2)
    assert n == 1, "push_back(ni) sites: %d" % n
    cc2, n = re.subn(r"lNodes\.push_front\((n[1-4])\);", r"\1.seq = orbxSeqCounter++; lNodes.push_front(\1);", cc2)
    assert n == 8, "push_front sites: %d" % n
    sort_re = r"sort\(vPrevSizeAndPointerToNode\.begin\(\),vPrevSizeAndPointerToNode\.end\(\)\);"
    cmp = ("sort(vPrevSizeAndPointerToNode.begin(),vPrevSizeAndPointerToNode.end(),"
           "[](const pair<int,ExtractorNode*>&a,const pair<int,ExtractorNode*>&b)"
           "{return a.first<b.first || (a.first==b.first && a.second->seq<b.second->seq);});")
    cc2, n = re.subn(sort_re, lambda m: cmp, cc2)
    assert n == 1, "careful-phase sort sites: %d" % n

    open(os.path.join(out, "ORBextractor.h"), "w").write(h2)
    open(os.path.join(out, "ORBextractor.cc"), "w").write(cc2)


if __name__ == "__main__":
    main()
