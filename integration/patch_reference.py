#!/usr/bin/env python
"""Generates the three reference files a maintainer would change to bind libgencore_b200.so from the
reference's own sources, at build time:

    python integration/patch_reference.py /root/reference/src  oracle/_ref/bridge_src

writes patched copies of gencore.h, gencore.cpp and reference.h (and main.cpp unchanged, so that it is compiled against the
patched gencore.h) into the output directory (git-ignored; nothing of the reference is stored in this repository).  Every edit is an exact-text replacement that must match exactly once, so a
different reference version fails loudly instead of producing a half-patched tree.  oracle/Makefile (target `bridge`)
compiles the result with integration/gcbbridge.h and the rest of the reference's sources into oracle/_ref/gencore_bridged.
"""
import os
import sys


def replace_once(text: str, old: str, new: str, what: str) -> str:
    n = text.count(old)
    if n != 1:
        raise SystemExit(f"patch_reference: anchor for '{what}' found {n} times (expected once) — not the reference version this was written for")
    return text.replace(old, new)


CALL_PROPER = """                vector<Pair*> csPairs = iter3->second->clusterByUMI(mOptions->properReadsUmiDiffThreshold, mPreStats, mPostStats, iter3->first < 0);
                for(int i=0; i<csPairs.size(); i++) {
                    //csPairs[i]->dump();
                    outputPair(csPairs[i]);
                    delete csPairs[i];
                }
                // this tid:left:right is done
                delete iter3->second;
"""
CALL_PROPER_NEW = """                // gencore_b200: the cluster joins the pending batch (the bridge owns it now); flushed after these loops
                bridge()->add(iter3->second, mOptions->properReadsUmiDiffThreshold, iter3->first < 0);
"""
WATERMARK = """    if(curProcessedTid != INT_MAX) {
        mProcessedTid = curProcessedTid;"""
WATERMARK_NEW = """    bridgeFlush();  // gencore_b200: one engine call for the clusters above, outputPair in the same order, before the watermark moves
    if(curProcessedTid != INT_MAX) {
        mProcessedTid = curProcessedTid;"""
CALL_UNPROPER = """                    vector<Pair*> csPairs = iter3->second->clusterByUMI(mOptions->unproperReadsUmiDiffThreshold, mPreStats, mPostStats, iter3->first < 0);
                    for(int i=0; i<csPairs.size(); i++) {
                        //csPairs[i]->dump();
                        outputPair(csPairs[i]);
                        delete csPairs[i];
                    }
                }
"""
CALL_UNPROPER_NEW = """                    // gencore_b200: see addToProperCluster
                    bridge()->add(iter3->second, mOptions->unproperReadsUmiDiffThreshold, iter3->first < 0);
                    iter3 = iter2->second.erase(iter3);
                    continue;
                }
"""
FINISH_TAIL = """        // this tid is done
        if(iter1->second.size() == 0) {
            iter1 = clusters.erase(iter1);
        } else {
            iter1++;
        }
    }
}
"""
FINISH_TAIL_NEW = """        // this tid is done
        if(iter1->second.size() == 0) {
            iter1 = clusters.erase(iter1);
        } else {
            iter1++;
        }
    }
    bridgeFlush();  // gencore_b200
}

// gencore_b200: the engine binding (integration/gcbbridge.h), created when the BAM header and the reference are known
GcbBridge* Gencore::bridge() {
    if(mBridge == NULL)
        mBridge = new GcbBridge(mOptions, mBamHeader);
    return mBridge;
}

void Gencore::bridgeFlush() {
    if(mBridge == NULL)
        return;
    mBridge->flush(mPreStats, mPostStats, [this](Pair* p) { outputPair(p); });
}
"""


def main() -> None:
    src, out = sys.argv[1], sys.argv[2]
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(src, "gencore.cpp")) as f:
        cpp = f.read()
    cpp = replace_once(cpp, '#include "gencore.h"\n', '#include "gencore.h"\n#include "gcbbridge.h"\n', "include")
    cpp = replace_once(cpp, "    mProperClustersFinished = false;\n}", "    mProperClustersFinished = false;\n    mBridge = NULL;\n}", "constructor")
    cpp = replace_once(cpp, "    delete mPreStats;\n    delete mPostStats;\n}", "    delete mBridge;\n    delete mPreStats;\n    delete mPostStats;\n}", "destructor")
    cpp = replace_once(cpp, CALL_PROPER, CALL_PROPER_NEW, "gencore.cpp:355")
    cpp = replace_once(cpp, WATERMARK, WATERMARK_NEW, "flush before the watermark")
    cpp = replace_once(cpp, CALL_UNPROPER, CALL_UNPROPER_NEW, "gencore.cpp:409")
    cpp = replace_once(cpp, FINISH_TAIL, FINISH_TAIL_NEW, "end of finishConsensus")
    with open(os.path.join(out, "gencore.cpp"), "w") as f:
        f.write(cpp)

    with open(os.path.join(src, "gencore.h")) as f:
        h = f.read()
    h = replace_once(h, "class Gencore {\n", "class GcbBridge;\n\nclass Gencore {\n", "forward declaration")
    h = replace_once(h, "    bool mProperClustersFinished;\n", "    bool mProperClustersFinished;\n    GcbBridge* mBridge;\n    GcbBridge* bridge();\n    void bridgeFlush();\n", "member")
    with open(os.path.join(out, "gencore.h"), "w") as f:
        f.write(h)

    with open(os.path.join(src, "reference.h")) as f:
        r = f.read()
    r = replace_once(r, "private:\n    FastaReader* mRef;", "    friend class GcbBridge;  // gencore_b200: reads the packed contigs\nprivate:\n    FastaReader* mRef;", "friend")
    with open(os.path.join(out, "reference.h"), "w") as f:
        f.write(r)
    # main.cpp is compiled from the same directory so that its #include "gencore.h" sees the class with the new member
    with open(os.path.join(src, "main.cpp")) as f:
        m = f.read()
    with open(os.path.join(out, "main.cpp"), "w") as f:
        f.write(m)
    print(f"patched gencore.cpp, gencore.h, reference.h (+ main.cpp as it is) -> {out}")


if __name__ == "__main__":
    main()
