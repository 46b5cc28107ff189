// GpuSortingCount.hpp -- host adapter: the SortingCountAlgorithm<span> surface that dsk's Functor<span> uses
// (R/src/DSK.cpp:45-70: ctor(bank, props), getInput(), execute(), getConfig(), getInfo(), getStorage(), getName()),
// implemented over the C ABI of include/dskgpu.h.  Written against gatb-core's own headers and linked with the
// reference's libgatbcore.a / libhdf5.a, so option parsing, bank opening and the HDF5 layout are the reference's own
// code; only the counting (fillPartitions + fillSolidKmers + the CountProcessor chain,
// K/SortingCountAlgorithm.cpp:636-781) is replaced by libdskgpu.so.  No CPU counting fallback exists here.
//
// Output contract, all written through gatb-core's Storage:
//   configuration.xml, minimizers/minimRepart, dsk/solid/<p> (+ attrs nb_partitions, kmer_size), dsk.xml,
//   histogram/{histogram,cutoff,nbsolidsforcutoff}, <out>.histo / <out>.histo2D text files.
#pragma once

#include <gatb/gatb_core.hpp>
#include <gatb/kmer/impl/ConfigurationAlgorithm.hpp>
#include <gatb/kmer/impl/CountProcessorHistogram.hpp>
#include <zlib.h>

#include <sstream>
#include <string>
#include <vector>

#include "../include/dskgpu.h"

namespace dskgpu_host {

using namespace gatb::core;
using namespace gatb::core::system;
using namespace gatb::core::system::impl;
using namespace gatb::core::bank;
using namespace gatb::core::bank::impl;
using namespace gatb::core::kmer;
using namespace gatb::core::kmer::impl;
using namespace gatb::core::tools::misc;
using namespace gatb::core::tools::misc::impl;
using namespace gatb::core::tools::storage::impl;
using namespace gatb::core::tools::collections;
using namespace gatb::core::tools::dp;

/** Errors cross the C ABI as codes; above it they are the reference's exception type (R/src/main.cpp:37-47). */
inline void check(int rc, dskgpu_ctx* ctx, const char* what)
{
    if (rc == DSKGPU_OK) return;
    const char* detail = dskgpu_last_error(ctx);
    throw Exception("%s: %s (%s)", what, dskgpu_strerror(rc), detail ? detail : "");
}

/** Thrown instead of Exception when the device record scanner rejects the input layout (DSKGPU_ERR_FORMAT): the adapter
 *  answers by feeding the same banks through the reference's own parser (IBank::iterator), see execute(). */
struct FormatRejected { std::string what; };
inline void checkFormat(int rc, dskgpu_ctx* ctx, const char* what)
{
    if (rc == DSKGPU_ERR_FORMAT) { FormatRejected e; const char* d = dskgpu_last_error(ctx); e.what = d ? d : ""; throw e; }
    check(rc, ctx, what);
}

template <size_t span = KMER_DEFAULT_SPAN>
class GpuSortingCount : public Algorithm
{
public:
    typedef typename Kmer<span>::Type  Type;
    typedef typename Kmer<span>::Count Count;

    /** Same meaning as SortingCountAlgorithm(IBank*, IProperties*) (K/SortingCountAlgorithm.cpp:119-133). */
    GpuSortingCount(IBank* bank, IProperties* params)
        : Algorithm("dsk", -1, params), _bank(0), _storage(0), _ctx(0), _solidCounts(0)
    {
        setBank(bank);
    }

    ~GpuSortingCount()
    {
        if (_ctx) dskgpu_destroy(_ctx);
        setBank(0);
        setStorage(0);
    }

    /** Same option parser as the reference (static, library code reused as is). */
    static IOptionsParser* getOptionsParser(bool mandatory = true) { return SortingCountAlgorithm<span>::getOptionsParser(mandatory); }

    const Configuration& getConfig() const { return _config; }
    Storage*             getStorage()      { return _storage; }
    Partition<Count>*    getSolidCounts()  { return _solidCounts; }

    void execute()
    {
        configure();
        // The device scanner takes what dsk is fed in practice: FASTA (single- or multi-line) and 4-line FASTQ, plain or
        // gzip.  Whatever else BankFasta accepts (multi-line FASTQ, '+' lines inside FASTA, ...; BankFasta.cpp:485-572) is
        // rejected by the scanner, never mis-parsed; the same banks then go through the reference's own parser, sequence by
        // sequence (IBank::iterator -> dskgpu_push_reads), and the counting path is unchanged.
        for (int attempt = 0; ; attempt++) {
            try {
                {
                    TIME_INFO(getTimeInfo(), "fill_partitions");          // same labels as K/SortingCountAlgorithm.cpp:1218
                    feedBanks();
                }
                {
                    TIME_INFO(getTimeInfo(), "fill_solid_kmers");         // K/SortingCountAlgorithm.cpp:1391
                    checkFormat(dskgpu_finish(_ctx), _ctx, "dskgpu_finish");
                    if (_autoCutoff) executeAutoCutoff();
                }
                break;
            } catch (FormatRejected& e) {
                if (attempt > 0 || _viaIterator) throw Exception("dskgpu: input rejected by the record scanner (%s)", e.what.c_str());
                _viaIterator = true;
                check(dskgpu_reset(_ctx), _ctx, "dskgpu_reset");
            }
        }
        writeResults();
    }

private:
    IBank* _bank;
    void setBank(IBank* bank) { SP_SETATTR(bank); }
    Storage* _storage;
    void setStorage(Storage* storage) { SP_SETATTR(storage); }

    Configuration     _config;
    bool              _viaIterator = false;        // second attempt: banks parsed by the reference's own reader
    dskgpu_ctx*       _ctx;
    Partition<Count>* _solidCounts;
    dskgpu_stats      _st;
    std::string       _histoName, _histo2DName;
    bool              _autoCutoff;
    bool              _autoPerBank;
    std::vector<long long> _userAbundanceMin;      // -1 = auto
    std::vector<CountNumber> _cutoffs;

    // ---- configure(): K/SortingCountAlgorithm.cpp:525-625 -------------------------------------------------------
    void configure()
    {
        IProperties* in = getInput();
        if (_bank == 0) { setBank(Bank::open(in->getStr(STR_URI_INPUT))); }

        std::string output = in->get(STR_URI_OUTPUT) ? in->getStr(STR_URI_OUTPUT)
                                                     : (in->getStr(STR_URI_OUTPUT_DIR) + "/" + System::file().getBaseName(_bank->getId()));
        if (!System::file().doesExist(in->getStr(STR_URI_OUTPUT_DIR))) {
            if (System::file().mkdir(in->getStr(STR_URI_OUTPUT_DIR), 0755) != 0) throw Exception("Error: can't create output directory");
        }
        std::string storage_type = in->getStr(STR_STORAGE_TYPE);
        StorageMode_e mode;
        if (storage_type == "hdf5") mode = STORAGE_HDF5;
        else if (storage_type == "file") mode = STORAGE_FILE;
        else throw Exception("Error: unknown storage type specified: %s", storage_type.c_str());
        setStorage(StorageFactory(mode).create(output, true, false));

        // the reference's own configuration step: parses thresholds / solidity kind / banks, estimates the volume
        ConfigurationAlgorithm<span> configAlgo(_bank, in);
        configAlgo.execute();
        _config = configAlgo.getConfiguration();
        // passes / partitions are a property of the device path: no disk tier, one ordered output collection
        _config._nb_passes = 1;
        _config._nb_partitions = 1;
        _storage->getGroup(configAlgo.getName()).setProperty("xml", std::string("\n") + configAlgo.getInfo()->getXML());

        const bool histo2D = in->get(STR_HISTO2D) && in->getInt(STR_HISTO2D) != 0;
        const bool histo1D = in->get(STR_HISTO) && in->getInt(STR_HISTO) != 0;
        if (histo2D) {                                              // K/SortingCountAlgorithm.cpp:604-620
            if (_bank->getBanks().size() < 2) throw Exception("There must be at least 2 input banks when using -histo2D");
        }
        // file names as in getDefaultProcessor (K/SortingCountAlgorithm.cpp:269-330)
        std::string base;
        if (in->get(STR_URI_OUTPUT)) base = in->getStr(STR_URI_OUTPUT);
        else {
            std::string uri = in->get(STR_URI_INPUT) ? in->getStr(STR_URI_INPUT) : (in->get(STR_URI_FILE) ? in->getStr(STR_URI_FILE) : std::string("resultfile"));
            base = System::file().getBaseName(uri.substr(0, uri.find(",")));
        }
        _histoName = histo1D ? base + ".histo" : std::string();
        _histo2DName = histo2D ? base + ".histo2D" : std::string();

        if (in->getInt(STR_HISTOGRAM_MAX) != 10000)
            throw Exception("the device histogram has the reference's default length (-histo-max 10000) only");
        if (in->getInt(STR_MINIMIZER_TYPE) != 0 || in->getInt(STR_REPARTITION_TYPE) != 0)
            throw Exception("-minimizer-type 1 / -repartition-type 1 are outside the device path (SURVEY.md 8(f)-4)");

        // ---- device context ---------------------------------------------------------------------------------------
        dskgpu_config c;
        dskgpu_config_default(&c);
        c.kmer_size = (int32_t)_config._kmerSize;
        // -minimizer-size left at the reference default: sized from the estimated volume, the way ConfigurationAlgorithm
        // sizes nb_partitions (bins must shrink as the job grows; which partition a k-mer lands in is unobservable)
        c.minimizer_size = (int32_t)in->getInt(STR_MINIMIZER_SIZE);
        if (c.minimizer_size == 10) c.minimizer_size = dskgpu_suggest_minimizer_size((uint64_t)_config._kmersNb, c.kmer_size);
        c.nb_banks = (int32_t)_config._nb_banks;
        if (c.nb_banks > DSKGPU_MAX_BANKS) throw Exception("at most %d banks are supported by the device path", DSKGPU_MAX_BANKS);
        // kind of the FILTER: created from -solidity-kind before -histo2D forces the counting path to per-bank
        // (K/SortingCountAlgorithm.cpp:600 vs :608-609) -- reproduced on purpose
        int kind = DSKGPU_SOLIDITY_SUM;
        switch (_config._solidityKind) {
        case KMER_SOLIDITY_MIN: kind = DSKGPU_SOLIDITY_MIN; break;
        case KMER_SOLIDITY_MAX: kind = DSKGPU_SOLIDITY_MAX; break;
        case KMER_SOLIDITY_ONE: kind = DSKGPU_SOLIDITY_ONE; break;
        case KMER_SOLIDITY_ALL: kind = DSKGPU_SOLIDITY_ALL; break;
        case KMER_SOLIDITY_CUSTOM: kind = DSKGPU_SOLIDITY_CUSTOM; break;
        default: kind = DSKGPU_SOLIDITY_SUM; break;
        }
        if (c.nb_banks == 1) kind = DSKGPU_SOLIDITY_SUM;           // K/ConfigurationAlgorithm.cpp:261-264
        c.solidity_kind = kind;
        c.per_bank_counts = (c.nb_banks > 1 && (kind != DSKGPU_SOLIDITY_SUM || histo2D)) ? 1 : 0;
        c.histo2d = histo2D ? 1 : 0;
        _autoCutoff = false;
        _userAbundanceMin.clear();
        for (size_t i = 0; i < (size_t)DSKGPU_MAX_BANKS; i++) {
            const size_t j = i < _config._abundance.size() ? i : _config._abundance.size() - 1;
            long long lo = _config._abundance.empty() ? 2 : (long long)_config._abundance[j].getBegin();
            if (lo < 0) { _autoCutoff = true; lo = -1; }
            _userAbundanceMin.push_back(lo);
            c.abundance_min[i] = lo;
        }
        // "auto" anywhere: the reference runs the partitions through a cutoff processor first, then through the dsk
        // chain (K/SortingCountAlgorithm.cpp:455-514).  Here: pass 1 dumps nothing (no abundance reaches the threshold)
        // and leaves the histogram(s); executeAutoCutoff() turns them into thresholds and counts again from HBM.
        _autoPerBank = _autoCutoff && c.nb_banks > 1 && (kind == DSKGPU_SOLIDITY_ONE || kind == DSKGPU_SOLIDITY_ALL || kind == DSKGPU_SOLIDITY_CUSTOM);
        if (_autoCutoff) for (size_t i = 0; i < (size_t)DSKGPU_MAX_BANKS; i++) c.abundance_min[i] = 2147483647LL;
        c.bank_histograms = _autoPerBank ? 1 : 0;
        c.abundance_max = _config._abundance.empty() ? 2147483647LL : (long long)_config._abundance[0].getEnd();
        for (size_t i = 0; i < (size_t)DSKGPU_MAX_BANKS; i++) c.solid_vec[i] = (i < _config._solidVec.size()) ? (_config._solidVec[i] ? 1 : 0) : 1;
        const char* dev = getenv("DSKGPU_DEVICE");
        c.device = dev ? atoi(dev) : 0;
        check(dskgpu_create(&c, &_ctx), 0, "dskgpu_create");
    }

    // ---- fillPartitions(): K/SortingCountAlgorithm.cpp:1216-1349, replaced by streaming file bytes to the device ----
    static bool isRegularFile(const std::string& p) { return !p.empty() && System::file().doesExist(p) && System::file().getSize(p) > 0; }

    void feedFile(int bankId, const std::string& path, char* buf[2], size_t cap)
    {
        // zlib reads plain and gzip files alike (what BankFasta does, G/src/gatb/bank/impl/BankFasta.cpp:391-396)
        gzFile f = gzopen(path.c_str(), "rb");
        if (!f) throw Exception("unable to open file %s", path.c_str());
        gzbuffer(f, 1 << 20);
        int par = 0;
        size_t n = 0;
        int r = gzread(f, buf[par], (unsigned)cap);
        if (r < 0) { gzclose(f); throw Exception("read error on %s", path.c_str()); }
        n = (size_t)r;
        for (;;) {
            // read the next block before pushing this one, so the last block is known to be the last
            int r2 = (n == cap) ? gzread(f, buf[par ^ 1], (unsigned)cap) : 0;
            if (r2 < 0) { gzclose(f); throw Exception("read error on %s", path.c_str()); }
            const bool last = (r2 == 0);
            const int rc = dskgpu_push_bytes(_ctx, bankId, buf[par], n, DSKGPU_FMT_AUTO, last ? DSKGPU_PUSH_LAST : 0);
            if (rc != DSKGPU_OK) { gzclose(f); checkFormat(rc, _ctx, "dskgpu_push_bytes"); }
            if (last) break;
            par ^= 1; n = (size_t)r2;
        }
        gzclose(f);
    }

    void feedSequences(int bankId, IBank* b)
    {
        // non-file banks (BankStrings, BankRandom, ...): IBank::iterator() -> concatenated sequences
        std::vector<char> bases; std::vector<uint64_t> offs(1, 0);
        Iterator<Sequence>* it = b->iterator(); LOCAL(it);
        for (it->first(); !it->isDone(); it->next()) {
            Sequence& s = it->item();
            bases.insert(bases.end(), s.getDataBuffer(), s.getDataBuffer() + s.getDataSize());
            offs.push_back(bases.size());
            if (bases.size() > ((size_t)256 << 20)) {
                check(dskgpu_push_reads(_ctx, bankId, bases.data(), offs.data(), offs.size() - 1), _ctx, "dskgpu_push_reads");
                bases.clear(); offs.assign(1, 0);
            }
        }
        if (offs.size() > 1) check(dskgpu_push_reads(_ctx, bankId, bases.data(), offs.data(), offs.size() - 1), _ctx, "dskgpu_push_reads");
    }

    void collectLeaves(IBank* b, std::vector<IBank*>& out)
    {
        const std::vector<IBank*> sub = b->getBanks();
        if (sub.size() == 1 && sub[0] == b) { out.push_back(b); return; }
        if (sub.empty()) { out.push_back(b); return; }
        for (size_t i = 0; i < sub.size(); i++) collectLeaves(sub[i], out);
    }

    void feedBanks()
    {
        const size_t cap = (size_t)64 << 20;
        char* buf[2] = {(char*)dskgpu_host_alloc(cap), (char*)dskgpu_host_alloc(cap)};
        if (!buf[0] || !buf[1]) throw Exception("pinned staging allocation failed");
        try {
            // bank id = index in the top-level composition, as the reference's per-bank counts (getCompositionNb)
            const std::vector<IBank*> top = _bank->getBanks();
            const bool composite = _config._nb_banks > 1 && top.size() == _config._nb_banks;
            for (size_t t = 0; t < (composite ? top.size() : 1); t++) {
                std::vector<IBank*> leaves;
                collectLeaves(composite ? top[t] : _bank, leaves);
                for (size_t i = 0; i < leaves.size(); i++) {
                    const std::string id = leaves[i]->getId();
                    if (!_viaIterator && isRegularFile(id)) feedFile((int)t, id, buf, cap);
                    else feedSequences((int)t, leaves[i]);
                }
            }
        } catch (...) { dskgpu_host_free(buf[0]); dskgpu_host_free(buf[1]); throw; }
        dskgpu_host_free(buf[0]); dskgpu_host_free(buf[1]);
    }

    // ---- -abundance-min auto: CountProcessorCutoff::endPass + CountProcessorSolidityInfo::setAbundanceMin ----------------
    // (K/CountProcessorCutoff.hpp:86-124, K/CountProcessorSolidity.hpp:45-66); the heuristic itself is the reference's own
    // Histogram::compute_threshold, run on the device histogram(s)
    void executeAutoCutoff()
    {
        const size_t nbHist = _autoPerBank ? (size_t)_config._nb_banks : 1;
        std::vector<uint64_t> hs(nbHist * (size_t)DSKGPU_HISTO_LEN);
        if (_autoPerBank) check(dskgpu_bank_histograms(_ctx, hs.data()), _ctx, "dskgpu_bank_histograms");
        else              check(dskgpu_histogram(_ctx, hs.data(), 0), _ctx, "dskgpu_histogram");
        _cutoffs.clear();
        for (size_t b = 0; b < nbHist; b++) {
            Histogram H(10000);
            for (size_t i = 0; i <= 10000; i++) H.get((u_int16_t)i) = hs[b * DSKGPU_HISTO_LEN + i];
            H.compute_threshold(3);
            _cutoffs.push_back((CountNumber)H.get_solid_cutoff());
        }
        const size_t nbThr = std::max<size_t>(1, _config._abundance.size());
        if (_cutoffs.size() > nbThr) throw Exception("Unable to set abundance min values (%d values for %d banks)", (int)_cutoffs.size(), (int)nbThr);
        int64_t amin[DSKGPU_MAX_BANKS];
        for (size_t i = 0; i < _cutoffs.size(); i++) amin[i] = _userAbundanceMin[i] == -1 ? (int64_t)_cutoffs[i] : (int64_t)_userAbundanceMin[i];
        for (size_t i = _cutoffs.size(); i < (size_t)DSKGPU_MAX_BANKS; i++) amin[i] = amin[_cutoffs.size() - 1];
        check(dskgpu_recount(_ctx, amin), _ctx, "dskgpu_recount");
    }

    // ---- results: CountProcessorDump / CountProcessorHistogram roles, in bulk ---------------------------------------
    void writeResults()
    {
        IProperties* in = getInput();
        check(dskgpu_get_stats(_ctx, &_st), _ctx, "dskgpu_get_stats");

        // minimizers/minimRepart: the blob of Repartitor::save (K/PartiInfo.cpp:271-295) for our map: every minimizer
        // goes to the single output collection 0
        {
            Group& g = _storage->getGroup("minimizers");
            const u_int16_t nbpart = 1, nbpass = 1;
            const u_int64_t nbMinims = (u_int64_t)1 << (2 * _config._minim_size);
            const std::vector<u_int16_t> table(nbMinims, 0);
            const bool hasFreq = false;
            const u_int32_t magic = 0x12345678;
            Storage::ostream os(g, "minimRepart");
            os.write((const char*)&nbpart, sizeof nbpart);
            os.write((const char*)&nbMinims, sizeof nbMinims);
            os.write((const char*)&nbpass, sizeof nbpass);
            os.write((const char*)table.data(), sizeof(u_int16_t) * nbMinims);
            os.write((const char*)&hasFreq, sizeof hasFreq);
            os.write((const char*)&magic, sizeof magic);
            os.flush();
        }

        // dsk/solid/<p>: CountProcessorDump::begin + bulk Bag<Count>::insert (K/CountProcessorDump.hpp:85-152)
        Group& dsk = _storage->getGroup("dsk");
        const int nparts = dskgpu_num_partitions(_ctx);
        if (nparts < 0) check(nparts, _ctx, "dskgpu_num_partitions");
        _solidCounts = &dsk.template getPartition<Count>("solid", (size_t)nparts);
        dsk.addProperty("kmer_size", Stringify::format("%d", (int)_config._kmerSize));
        u_int64_t nbSolid = 0;
        std::vector<Count> block;
        for (int p = 0; p < nparts; p++) {
            const uint64_t* kmers = 0; const uint32_t* counts = 0; uint64_t n = 0; int words = 0;
            check(dskgpu_partition(_ctx, p, &kmers, &counts, &n, &words), _ctx, "dskgpu_partition");
            const size_t B = 1 << 20;
            for (uint64_t i = 0; i < n; i += B) {
                const size_t m = (size_t)std::min<uint64_t>(B, n - i);
                block.resize(m);
                for (size_t j = 0; j < m; j++) {
                    Type v; setValue(v, kmers + (i + j) * words, words);
                    block[j] = Count(v, (CountNumber)counts[i + j]);
                }
                (*_solidCounts)[p].insert(block.data(), m);
            }
            (*_solidCounts)[p].flush();
            nbSolid += n;
        }
        _solidCounts->flush();

        // histogram group + text files: the reference's own CountProcessorHistogram::end() on a Histogram object that
        // was filled from the device bins (K/CountProcessorHistogram.hpp:104-159, Histogram.cpp:43-190)
        std::vector<uint64_t> h1(DSKGPU_HISTO_LEN), h2((size_t)DSKGPU_HISTO_LEN * DSKGPU_HISTO2D_DIM2);
        check(dskgpu_histogram(_ctx, h1.data(), h2.data()), _ctx, "dskgpu_histogram");
        const bool histo2D = !_histo2DName.empty(), histo1D = !_histoName.empty();
        CountProcessorHistogram<span> ph(&_storage->getGroup("histogram"), 10000, in->getInt(STR_KMER_ABUNDANCE_MIN_THRESHOLD),
                                         histo2D, histo1D, _histo2DName, _histoName);
        IHistogram* H = ph.getHistogram();
        for (size_t i = 0; i <= 10000; i++) H->get((u_int16_t)i) = h1[i];
        for (size_t j = 0; j <= 10; j++) for (size_t i = 0; i <= 10000; i++) H->get2D((u_int16_t)i, (u_int16_t)j) = h2[j * DSKGPU_HISTO_LEN + i];
        ph.end();

        // ---- statistics: keys of K/SortingCountAlgorithm.cpp:728-780 ----------------------------------------------------
        getInfo()->add(1, "bank");
        getInfo()->add(2, "bank_uri", "%s", _bank->getId().c_str());
        getInfo()->add(2, "bank_size", "%lld", (long long)_bank->getSize());
        getInfo()->add(2, "bank_total_nt", "%lld", (long long)_st.nb_nucleotides);
        getInfo()->add(2, "sequences");
        getInfo()->add(3, "seq_number", "%ld", (long)_st.nb_sequences);
        getInfo()->add(3, "seq_size_mean", "%.1f", _st.nb_sequences ? (double)_st.nb_nucleotides / (double)_st.nb_sequences : 0.0);
        getInfo()->add(2, "kmers");
        getInfo()->add(3, "kmers_nb_valid", "%lld", (long long)_st.kmers_nb_valid);
        getInfo()->add(1, "stats");
        getInfo()->add(2, "temp_files");
        getInfo()->add(3, "nb_superkmers", "%lld", (long long)_st.nb_superkmers);
        getInfo()->add(3, "avg_superk_length", "%.2f", _st.nb_superkmers ? (double)_st.kmers_nb_valid / (double)_st.nb_superkmers : 0.0);
        getInfo()->add(3, "total_size_(MB)", "%lld", (long long)(_st.superkmer_bytes >> 20));
        if (_autoCutoff) {                                        // CountProcessorCutoff::getProperties (K/CountProcessorCutoff.hpp:128-137)
            std::stringstream ss; for (size_t i = 0; i < _cutoffs.size(); i++) ss << _cutoffs[i] << " ";
            getInfo()->add(2, "cutoffs_auto");
            getInfo()->add(3, "values", "%s", ss.str().c_str());
        }
        getInfo()->add(2, ph.getProperties());
        getInfo()->add(2, "kmers");
        getInfo()->add(3, "solidity_kind", "%s", toString(_config._solidityKind).c_str());
        getInfo()->add(3, "kmers_nb_distinct", "%ld", (long)_st.kmers_nb_distinct);
        getInfo()->add(3, "kmers_nb_solid", "%ld", (long)nbSolid);
        getInfo()->add(3, "kmers_nb_weak", "%ld", (long)(_st.kmers_nb_distinct - nbSolid));
        if (_st.kmers_nb_distinct) getInfo()->add(3, "kmers_percent_weak", "%.1f", 100.0 - 100.0 * (double)nbSolid / (double)_st.kmers_nb_distinct);
        getInfo()->add(2, "partitions");
        getInfo()->add(3, "nb_partitions", "%ld", (long)nparts);
        getInfo()->add(3, "nb_items", "%ld", (long)nbSolid);
        getInfo()->add(3, "device_partitions", "%ld", (long)_st.nb_partitions);
        getInfo()->add(3, "kind");
        getInfo()->add(4, "vector", "%ld", (long)_st.nb_groups_sort);
        getInfo()->add(4, "hash", "%ld", (long)_st.nb_groups_hash);
        getInfo()->add(2, "device_time_ms");
        getInfo()->add(3, "parse", "%.3f", _st.ms_parse);
        getInfo()->add(3, "superkmers", "%.3f", _st.ms_superk);
        getInfo()->add(3, "partition", "%.3f", _st.ms_partition);
        getInfo()->add(3, "count", "%.3f", _st.ms_count);
        getInfo()->add(3, "order", "%.3f", _st.ms_sort);
        getInfo()->add(3, "gpu_launches", "%ld", (long)_st.gpu_launches);
        getInfo()->add(1, getTimeInfo().getProperties("time"));
    }

    // Type = LargeInt<1> (u64) for span 32, LargeInt<2> (__uint128_t or 2 x u64) for span 64: both expose their
    // words through getVal()/setVal or operator[]; going through shifts keeps this independent of the representation
    static void setValue(Type& v, const uint64_t* w, int words)
    {
        v.setVal(w[words - 1]);
        for (int i = words - 2; i >= 0; i--) { v <<= 32; v <<= 32; Type lo; lo.setVal(w[i]); v = v + lo; }
    }
};

}  // namespace dskgpu_host
