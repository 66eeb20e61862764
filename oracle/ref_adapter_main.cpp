// TEST INFRASTRUCTURE ONLY.  Drives the reference-tree binding of the CUDA path
// (tsxcount_b200/host/ref_binding/TSXHashMapCUDA_ref.h) through the REFERENCE's own types: k-mers are made by
// TSXSeqUtils::fromSequence (src/utils/SequenceUtils.h:86), added through the virtual TSXHashMap::addKmer and read
// back through getKmerCount(kmer).toUInt() exactly like evaluate() does (src/mains/testExecution.h:38-50).
// Compiled here against /root/reference by oracle/Makefile; run on the GPU box by tests/test_cli.py.
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include <src/tsxcount/TSXHashMap.h>
#include <src/utils/SequenceUtils.h>
#include <src/mains/testExecution.h>

#include "TSXHashMapCUDA_ref.h"

int main(int argc, char** argv) {
    const bool compile_only = argc > 1 && std::string(argv[1]) == "--no-gpu";
    if (compile_only) { std::cout << "ref-adapter built" << std::endl; return 0; }
    const uint16_t k = 14;
    TSXHashMap* pMap = new TSXHashMapCUDA(20, 4, k, 1);            // through the base-class pointer, like main.cpp:429-475
    const std::string read = "ATCGAGTCAGTAGGCTTAACCGTTAGACCAGTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT";
    std::vector<std::string> kmers = createKMers(const_cast<std::string&>(read), k, pMap->getMemoryPool());   // testExecution.h:15-36
    std::map<std::string, uint64_t> want;
    for (int rep = 0; rep < 3; ++rep)
        for (auto& s : kmers) {
            TSX::tsx_kmer_t oKmer = TSXSeqUtils::fromSequence(s, pMap->getMemoryPool());
            pMap->addKmer(oKmer);
            ++want[s];
        }
    // the same read once more through the batch path: counts must add up
    static_cast<TSXHashMapCUDA*>(pMap)->addReads({read});
    for (auto& kv : want) kv.second += kv.second / 3;
    int errors = 0;
    for (auto& kv : want) {
        std::string s = kv.first;
        TSX::tsx_kmer_t oKmer = TSXSeqUtils::fromSequence(s, pMap->getMemoryPool());
        const uint64_t got = pMap->getKmerCount(oKmer).toUInt();
        if (got != kv.second) { ++errors; std::cerr << s << " expected " << kv.second << " got " << got << std::endl; }
    }
    const uint64_t distinct = static_cast<TSXHashMapCUDA*>(pMap)->getKmerCount();
    if (distinct != want.size()) { ++errors; std::cerr << "distinct " << distinct << " != " << want.size() << std::endl; }
    std::cout << "total errors" << errors << " distinct " << distinct << std::endl;    // the reference's own wording, main.cpp:369
    delete pMap;
    return errors ? 1 : 0;
}
