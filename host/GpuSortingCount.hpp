// GpuSortingCount.hpp -- host adapter: the SortingCountAlgorithm<span> surface that dsk's Functor<span> uses
// (R/src/DSK.cpp:45-70: ctor(bank, props), getInput(), execute(), getConfig(), getInfo(), getStorage(), getName()),
// implemented over the C ABI of include/dskgpu.h.  Written against gatb-core's own headers and linked with the
// reference's libgatbcore.a / libhdf5.a, so option parsing, bank opening and the HDF5 layout are the reference's own
// code; only the counting (fillPartitions + fillSolidKmers + the CountProcessor chain,
// K/SortingCountAlgorithm.cpp:636-781) is replaced by libdskgpu.so.  No CPU counting fallback exists here.
//
// Several GPUs (DSKGPU_DEVICES=all | n | "0,2,5"): ONE process, one context + one reader thread per device.  Plain files are
// cut into byte ranges at record starts and every device parses its own range; dskgpu_multi_finish routes the super-k-mers to
// the device owning their partition (peer stores) and every device counts what it owns -- the role the reference gives to
// its partition files and worker threads.  Jobs whose records exceed HBM run as several passes over the input
// (DSKGPU_NB_PASSES or dskgpu_suggest_nb_passes; the reference's pass loop, K/SortingCountAlgorithm.cpp:678-689).
// Output collections: dsk/solid/<pass * nb_devices + device>, k-mers ascending inside each one.
//
// Output contract, all written through gatb-core's Storage:
//   configuration.xml, minimizers/minimRepart, dsk/solid/<p> (+ attrs nb_partitions, kmer_size), dsk.xml,
//   histogram/{histogram,cutoff,nbsolidsforcutoff}, <out>.histo / <out>.histo2D text files.
#pragma once

#include <gatb/gatb_core.hpp>
#include <gatb/kmer/impl/ConfigurationAlgorithm.hpp>
#include <gatb/kmer/impl/CountProcessorHistogram.hpp>
#include <zlib.h>
#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
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

/** DSKGPU_TRACE=1: host wall-clock of the adapter's phases on stderr, next to libdskgpu's own trace (tuning aid). */
inline void hostTrace(const char* what)
{
    static const bool on = getenv("DSKGPU_TRACE") != 0;
    static const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    if (on) fprintf(stderr, "[dsk_gpu host] %-44s @%9.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
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
    typedef ICountProcessor<span>      CountProcessor;

    /** Same meaning as SortingCountAlgorithm(IBank*, IProperties*) (K/SortingCountAlgorithm.cpp:119-133). */
    GpuSortingCount(IBank* bank, IProperties* params)
        : Algorithm("dsk", -1, params), _bank(0), _storage(0), _ctx(0), _solidCounts(0), _nbPasses(1), _repartitor(0), _haveConfig(false)
    {
        setBank(bank);
        memset(&_st, 0, sizeof _st);
    }

    /** Same meaning as SortingCountAlgorithm(bank, config, repartitor, processors, params) (K/SortingCountAlgorithm.cpp:128-146,
     *  the constructor G/src/gatb/debruijn/impl/Graph.cpp:399-407 uses): the caller brings the Configuration and the count
     *  processors (ICountProcessor, G/src/gatb/kmer/api/ICountProcessor.hpp:105-173).  The repartitor is held and released
     *  but not consulted: the device path plans its own partitions from exact minimizer-bin counts, and which partition a
     *  k-mer lands in is unobservable to a count processor (it sees every distinct k-mer once, with its counts). */
    GpuSortingCount(IBank* bank, const Configuration& config, Repartitor* repartitor, std::vector<CountProcessor*> processors, IProperties* params)
        : Algorithm("dsk", config._nbCores, params), _bank(0), _storage(0), _config(config), _ctx(0), _solidCounts(0), _nbPasses(1),
          _repartitor(0), _haveConfig(true)
    {
        setBank(bank);
        setRepartitor(repartitor);
        memset(&_st, 0, sizeof _st);
        for (size_t i = 0; i < processors.size(); i++) addProcessor(processors[i]);
    }

    ~GpuSortingCount()
    {
        for (size_t i = 0; i < _ctxs.size(); i++) if (_ctxs[i]) dskgpu_destroy(_ctxs[i]);
        setBank(0);
        setStorage(0);
        setRepartitor(0);
        for (size_t i = 0; i < _processors.size(); i++) _processors[i]->forget();
    }

    /** The plug-in surface of SortingCountAlgorithm (K/SortingCountAlgorithm.hpp:151-166).  With at least one processor the
     *  device delivers EVERY distinct k-mer with its count and the processors decide what happens to it, exactly as in the
     *  reference's fillSolidKmers_aux (K/SortingCountAlgorithm.cpp:1391-1607); without any, the default chain (histogram ->
     *  solidity -> dump) runs fused on the device (the fast path `dsk_gpu` uses). */
    /** The caller saw -minimizer-size on the command line (IOptionsParser::saw): the value is used as given, even when it is
     *  the reference default, instead of being sized from the estimated volume (dskgpu_suggest_minimizer_size). */
    void            keepMinimizerSize(bool keep)     { _keepMinimizerSize = keep; }
    size_t          getProcessorNumber() const       { return _processors.size(); }
    CountProcessor* getProcessor(size_t idx)         { return _processors[idx]; }
    void            addProcessor(CountProcessor* p)  { p->use(); _processors.push_back(p); }

    /** Same option parser as the reference (static, library code reused as is). */
    static IOptionsParser* getOptionsParser(bool mandatory = true) { return SortingCountAlgorithm<span>::getOptionsParser(mandatory); }

    const Configuration& getConfig() const { return _config; }
    Storage*             getStorage()      { return _storage; }
    Partition<Count>*    getSolidCounts()  { return _solidCounts; }

    void execute()
    {
        hostTrace("execute: start");
        configure();
        hostTrace("configured (bank estimate, storage, device contexts)");
        if (!_processors.empty()) { executeWithProcessors(); return; }
        // The device scanner takes what dsk is fed in practice: FASTA (single- or multi-line) and 4-line FASTQ, plain or
        // gzip.  Whatever else BankFasta accepts (multi-line FASTQ, '+' lines inside FASTA, ...; BankFasta.cpp:485-572) is
        // rejected by the scanner, never mis-parsed; the same banks then go through the reference's own parser, sequence by
        // sequence (IBank::iterator -> dskgpu_push_reads), and the counting path is unchanged.
        for (int attempt = 0; ; attempt++) {
            try {
                runAllPasses();
                break;
            } catch (FormatRejected& e) {
                if (attempt > 0 || _viaIterator) throw Exception("dskgpu: input rejected by the record scanner (%s)", e.what.c_str());
                _viaIterator = true;
                for (size_t r = 0; r < _ctxs.size(); r++) check(dskgpu_reset(_ctxs[r]), _ctxs[r], "dskgpu_reset");
            }
        }
        hostTrace("all passes done (solid collections written)");
        writeResults();
        hostTrace("results written (histogram, minimRepart, stats)");
    }

private:
    IBank* _bank;
    void setBank(IBank* bank) { SP_SETATTR(bank); }
    Storage* _storage;
    void setStorage(Storage* storage) { SP_SETATTR(storage); }

    Configuration     _config;
    Repartitor*       _repartitor;
    void setRepartitor(Repartitor* repartitor) { SP_SETATTR(repartitor); }
    bool              _haveConfig;                 // the Configuration came with the constructor (5-argument form)
    std::vector<CountProcessor*> _processors;      // plug-in mode when not empty
    bool              _viaIterator = false;        // second attempt: banks parsed by the reference's own reader
    dskgpu_ctx*       _ctx;                        // context of rank 0 (== _ctxs[0])
    std::vector<dskgpu_ctx*> _ctxs;                // one context per device (rank r = _ctxs[r])
    int               _nbPasses;
    Partition<Count>* _solidCounts;
    dskgpu_stats      _st;                         // summed over the devices and the passes
    std::vector<uint64_t> _h1, _h2;                // histograms summed over the devices and the passes
    u_int64_t         _nbSolidWritten;
    std::string       _histoName, _histo2DName;
    size_t            _histoMax = 10000;
    int               _minimizerType = 0, _repartitionType = 0;
    bool              _keepMinimizerSize = false;
    int               _minimizerSizeUsed = 10;
    bool              _autoCutoff;
    bool              _autoPerBank;
    std::vector<long long> _userAbundanceMin;      // -1 = auto
    std::vector<CountNumber> _cutoffs;

    // ---- configure(): K/SortingCountAlgorithm.cpp:525-625 -------------------------------------------------------
    void configure()
    {
        IProperties* in = getInput();
        if (_bank == 0) { setBank(Bank::open(in->getStr(STR_URI_INPUT))); }

        bool plugin = !_processors.empty();
        // a default storage only when the caller did not bring processors of their own (K/SortingCountAlgorithm.cpp:534-575)
        if (!plugin || !_haveConfig) {
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
        }
        if (!_haveConfig) {
            // the reference's own configuration step: parses thresholds / solidity kind / banks, estimates the volume
            ConfigurationAlgorithm<span> configAlgo(_bank, in);
            configAlgo.execute();
            _config = configAlgo.getConfiguration();
            _storage->getGroup(configAlgo.getName()).setProperty("xml", std::string("\n") + configAlgo.getInfo()->getXML());
        }
        // -kff (KFF output next to the .h5): the reference's CountProcessorDumpKff sits in its default processor chain
        // (K/SortingCountAlgorithm.cpp:335-400).  The fused device chain has no KFF writer, so the job runs in plug-in mode with
        // the reference's own default processors -- histogram, solidity, dump and KFF dump are then all reference code fed by
        // the device.
        if (!plugin && in->get(STR_KFF)) {
            std::vector<CountProcessor*> dflt = SortingCountAlgorithm<span>::getDefaultProcessorVector(_config, in, _storage, _storage);
            for (size_t i = 0; i < dflt.size(); i++) addProcessor(dflt[i]);
            plugin = true;
        }
        // passes / partitions are a property of the device path (set below, once the devices are known): no disk tier, one
        // ordered output collection per device and pass
        _config._nb_passes = 1;
        _config._nb_partitions = 1;

        const bool histo2D = !plugin && in->get(STR_HISTO2D) && in->getInt(STR_HISTO2D) != 0;   // (plug-in mode: histograms are the processors' business)
        const bool histo1D = !plugin && in->get(STR_HISTO) && in->getInt(STR_HISTO) != 0;
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

        // -histo-max N (K/SortingCountAlgorithm.cpp:213): counts >= N fall into a clamp bin that is never merged
        // (Histogram.hpp:92,221), so a histogram of length N is the first N bins of the device's 10000-bin histogram
        _histoMax = (size_t)in->getInt(STR_HISTOGRAM_MAX);
        if (!plugin && (_histoMax < 1 || _histoMax > 10000))
            throw Exception("the device histogram holds 10000 bins: -histo-max %d is outside 1..10000", (int)_histoMax);
        // -minimizer-type 1 (frequency order, K/Model.hpp:957-973) and -repartition-type 1 (ordered repartition,
        // K/RepartitionAlgorithm.cpp:311-384) only change WHICH PARTITION a k-mer is counted in -- the reference uses them to
        // balance its partition files.  The device path balances its partitions itself, from exact minimizer-bin counts, and
        // delivers one ordered collection per device: both options are accepted and have nothing left to influence.
        _minimizerType = (int)in->getInt(STR_MINIMIZER_TYPE); _repartitionType = (int)in->getInt(STR_REPARTITION_TYPE);

        // ---- device context ---------------------------------------------------------------------------------------
        dskgpu_config c;
        dskgpu_config_default(&c);
        c.kmer_size = (int32_t)_config._kmerSize;
        c.sequence_stats = 1;                                      // BankStats (K/BankKmers.hpp:166-215): seq_size_min/max/deviation, kmers_nb_invalid
        // -minimizer-size left at the reference default: sized from the estimated volume, the way ConfigurationAlgorithm
        // sizes nb_partitions (bins must shrink as the job grows; which partition a k-mer lands in is unobservable)
        c.minimizer_size = (int32_t)in->getInt(STR_MINIMIZER_SIZE);
        if (c.minimizer_size == 10 && !_keepMinimizerSize) c.minimizer_size = dskgpu_suggest_minimizer_size((uint64_t)_config._kmersNb, c.kmer_size);
        _minimizerSizeUsed = c.minimizer_size;
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
        if (plugin) {
            // every distinct k-mer goes to the processors, with its count: nothing is filtered or histogrammed on the device.
            // A CountVector holds one count per bank (K/PartitionsCommand.cpp:540-541); the C ABI delivers one count per k-mer.
            // Several banks: one device job per bank, merged on the host into one CountVector per k-mer (executeWithProcessors).
            c.solidity_kind = DSKGPU_SOLIDITY_SUM; c.per_bank_counts = 0; c.histo2d = 0; c.bank_histograms = 0;
            for (size_t i = 0; i < (size_t)DSKGPU_MAX_BANKS; i++) c.abundance_min[i] = 1;
            c.abundance_max = 2147483647LL;
            _autoCutoff = false; _autoPerBank = false;
        }
        // devices: DSKGPU_DEVICES = "all" | a count | a comma list of ordinals; default one device (DSKGPU_DEVICE, 0)
        std::vector<int> devs;
        const char* dl = getenv("DSKGPU_DEVICES");
        if (dl && *dl) {
            const std::string v(dl);
            const int avail = dskgpu_device_count();
            if (v == "all") { for (int i = 0; i < avail; i++) devs.push_back(i); }
            else if (v.find(',') != std::string::npos) { std::stringstream ss(v); std::string tok; while (std::getline(ss, tok, ',')) if (!tok.empty()) devs.push_back(atoi(tok.c_str())); }
            else { const int n = atoi(v.c_str()); for (int i = 0; i < n; i++) devs.push_back(i); }
        }
        if (devs.empty()) { const char* dev = getenv("DSKGPU_DEVICE"); devs.push_back(dev ? atoi(dev) : 0); }
        const int W = (int)devs.size();
        // passes: the reference sizes them from the estimated volume against -max-disk (K/ConfigurationAlgorithm.cpp:245-467);
        // here the bound is what one pass keeps in HBM per device
        const char* np = getenv("DSKGPU_NB_PASSES");
        _nbPasses = np ? std::max(1, atoi(np)) : dskgpu_suggest_nb_passes((uint64_t)_config._kmersNb, c.kmer_size, W, 0, devs[0]);
        _config._nb_passes = (size_t)_nbPasses;
        _config._nb_partitions = (size_t)W;                        // output collections per pass: one per device, k-mers ascending
        c.world_size = W;
        c.nb_passes = _nbPasses; c.pass_id = 0;
        _ctxs.assign((size_t)W, (dskgpu_ctx*)0);
        for (int r = 0; r < W; r++) {
            c.rank = r; c.device = devs[(size_t)r];
            check(dskgpu_create(&c, &_ctxs[(size_t)r]), 0, "dskgpu_create");
        }
        _ctx = _ctxs[0];
    }

    // ---- fillPartitions(): K/SortingCountAlgorithm.cpp:1216-1349, replaced by streaming file bytes to the devices ----
    static bool isRegularFile(const std::string& p) { return !p.empty() && System::file().doesExist(p) && System::file().getSize(p) > 0; }

    /** What one reader thread feeds to its device: a byte range of a plain file, a whole (gzip) file, or a bank to iterate. */
    struct FeedTask { int bank; std::string path; uint64_t begin, end; bool whole; IBank* leaf; };
    /** An exception crossing a thread boundary (threads must not leak C++ exceptions). */
    struct ThreadError { bool failed, format; std::string what; ThreadError() : failed(false), format(false) {} };

    static bool isGzip(const std::string& path)
    {
        unsigned char h[2] = {0, 0};
        FILE* f = fopen(path.c_str(), "rb"); if (!f) return false;
        const size_t n = fread(h, 1, 2, f); fclose(f);
        return n == 2 && h[0] == 0x1f && h[1] == 0x8b;
    }

    /** First record start at or after `pos` in a plain FASTA / 4-line FASTQ file (returns `size` when there is none).
     *  FASTA: a line starting with '>'.  FASTQ: a line starting with '@' whose second next line starts with '+' (a quality
     *  line may start with '@', but then the line two below is a sequence line, which never starts with '+'). */
    static uint64_t nextRecordStart(int fd, uint64_t pos, uint64_t size, char first)
    {
        if (pos == 0) return 0;
        const size_t W = (size_t)4 << 20;
        std::vector<char> buf(W);
        uint64_t at = pos - 1;                                     // include the byte before: a record start follows a '\n'
        while (at < size) {
            const ssize_t n = pread(fd, buf.data(), W, (off_t)at);
            if (n <= 1) break;
            for (ssize_t i = 0; i + 1 < n; i++) {
                if (buf[(size_t)i] != '\n' || buf[(size_t)i + 1] != first) continue;
                if (first == '>') return at + (uint64_t)i + 1;
                // FASTQ: check the '+' line two below (inside this window; a window that ends first is retried from here)
                ssize_t e1 = i + 1; while (e1 < n && buf[(size_t)e1] != '\n') e1++;
                ssize_t e2 = e1 + 1; while (e2 < n && buf[(size_t)e2] != '\n') e2++;
                if (e2 + 1 >= n) { if (at + (uint64_t)n >= size) return size; goto refill; }
                if (buf[(size_t)e2 + 1] == '+') return at + (uint64_t)i + 1;
                continue;
            refill:
                at += i > 0 ? (uint64_t)i : (uint64_t)n - 1; goto next_window;        // (a record longer than the window cannot be a read)
            }
            at += (uint64_t)n - 1;
        next_window:;
        }
        return size;
    }

    /** Streams bytes [begin, end) of a plain file (or a whole gzip file through zlib, BankFasta.cpp:391-396) into one context.
     *  Two pinned buffers: the file read of block i+1 overlaps the device work on block i; dskgpu_push_sync guarantees the H2D
     *  copy out of a buffer has completed before that buffer is written again (push_bytes only QUEUES the copy). */
    static void feedRange(dskgpu_ctx* ctx, const FeedTask& t, char* buf[2], size_t cap)
    {
        gzFile gz = 0; int fd = -1;
        if (t.whole) { gz = gzopen(t.path.c_str(), "rb"); if (!gz) throw Exception("unable to open file %s", t.path.c_str()); gzbuffer(gz, 1 << 20); }
        else { fd = open(t.path.c_str(), O_RDONLY); if (fd < 0) throw Exception("unable to open file %s", t.path.c_str()); }
        uint64_t pos = t.begin;
        struct Closer { gzFile g; int f; ~Closer() { if (g) gzclose(g); if (f >= 0) close(f); } } closer = {gz, fd};
        (void)closer;
        auto readBlock = [&](char* dst) -> size_t {
            if (gz) { const int r = gzread(gz, dst, (unsigned)cap); if (r < 0) throw Exception("read error on %s", t.path.c_str()); return (size_t)r; }
            size_t got = 0;
            const size_t want = (size_t)std::min<uint64_t>(cap, t.end - pos);
            while (got < want) { const ssize_t r = pread(fd, dst + got, want - got, (off_t)(pos + got)); if (r < 0) throw Exception("read error on %s", t.path.c_str()); if (r == 0) break; got += (size_t)r; }
            pos += got;
            return got;
        };
        int par = 0;
        size_t n = readBlock(buf[par]);
        for (;;) {
            // read the next block before pushing this one, so the last block is known to be the last
            size_t n2 = 0;
            if (n == cap) { checkFormat(dskgpu_push_sync(ctx), ctx, "dskgpu_push_sync"); n2 = readBlock(buf[par ^ 1]); }
            const bool last = (n2 == 0);
            checkFormat(dskgpu_push_bytes(ctx, t.bank, buf[par], n, DSKGPU_FMT_AUTO, last ? DSKGPU_PUSH_LAST : 0), ctx, "dskgpu_push_bytes");
            if (last) break;
            par ^= 1; n = n2;
        }
        checkFormat(dskgpu_push_sync(ctx), ctx, "dskgpu_push_sync");       // the buffers are about to be freed / reused
    }

    /** Decompression ahead of the device: a gzip stream is serial (zlib, ~0.3-0.4 GB/s per stream), but the FILES of an album
     *  (paired-end R1 / R2, one file per lane, ...) are independent streams.  A device's task list is served in order by its
     *  reader thread while up to `ahead` helper threads inflate the next files into bounded queues of blocks: several zlib
     *  streams run at once, the device still sees every bank as one ordered byte stream.  (BankAlbum.cpp:310-348 reads its
     *  files one after the other on the calling thread.) */
    struct BlockQueue {
        std::mutex m; std::condition_variable cv; std::deque<std::vector<char> > q; bool done, cancel; std::string err;
        BlockQueue() : done(false), cancel(false) {}
    };
    static void inflateFile(const std::string& path, BlockQueue* Q, size_t blk, size_t depth)
    {
        gzFile gz = gzopen(path.c_str(), "rb");
        if (!gz) { std::unique_lock<std::mutex> lk(Q->m); Q->err = "unable to open file " + path; Q->done = true; Q->cv.notify_all(); return; }
        gzbuffer(gz, 1 << 20);
        for (;;) {
            std::vector<char> b(blk);
            const int r = gzread(gz, b.data(), (unsigned)blk);
            std::unique_lock<std::mutex> lk(Q->m);
            if (r < 0) { Q->err = "read error on " + path; break; }
            if (r == 0) break;
            b.resize((size_t)r);
            while (Q->q.size() >= depth && !Q->cancel) Q->cv.wait(lk);
            if (Q->cancel) break;
            Q->q.push_back(std::vector<char>()); Q->q.back().swap(b);
            Q->cv.notify_all();
            if ((size_t)r < blk) break;
        }
        gzclose(gz);
        std::unique_lock<std::mutex> lk(Q->m);
        Q->done = true; Q->cv.notify_all();
    }
    /** next block of a queue (empty vector = end of the file) */
    static std::vector<char> popBlock(BlockQueue* Q)
    {
        std::unique_lock<std::mutex> lk(Q->m);
        while (Q->q.empty() && !Q->done) Q->cv.wait(lk);
        if (!Q->err.empty()) throw Exception("%s", Q->err.c_str());
        std::vector<char> b;
        if (!Q->q.empty()) { b.swap(Q->q.front()); Q->q.pop_front(); Q->cv.notify_all(); }
        return b;
    }
    /** one whole file from its queue into the context: blocks come from pageable memory, which push_bytes may read
     *  synchronously (include/dskgpu.h); the block after the current one tells whether the current one is the last */
    static void feedQueue(dskgpu_ctx* ctx, const FeedTask& t, BlockQueue* Q)
    {
        std::vector<char> cur = popBlock(Q);
        for (;;) {
            std::vector<char> next;
            if (!cur.empty()) next = popBlock(Q);
            const bool last = next.empty();
            checkFormat(dskgpu_push_bytes(ctx, t.bank, cur.empty() ? "" : cur.data(), cur.size(), DSKGPU_FMT_AUTO, last ? DSKGPU_PUSH_LAST : 0), ctx, "dskgpu_push_bytes");
            if (last) break;
            cur.swap(next);
        }
        checkFormat(dskgpu_push_sync(ctx), ctx, "dskgpu_push_sync");
    }

    static void feedSequences(dskgpu_ctx* ctx, int bankId, IBank* b)
    {
        // non-file banks (BankStrings, BankRandom, ...): IBank::iterator() -> concatenated sequences
        std::vector<char> bases; std::vector<uint64_t> offs(1, 0);
        Iterator<Sequence>* it = b->iterator(); LOCAL(it);
        for (it->first(); !it->isDone(); it->next()) {
            Sequence& s = it->item();
            bases.insert(bases.end(), s.getDataBuffer(), s.getDataBuffer() + s.getDataSize());
            offs.push_back(bases.size());
            if (bases.size() > ((size_t)256 << 20)) {
                check(dskgpu_push_reads(ctx, bankId, bases.data(), offs.data(), offs.size() - 1), ctx, "dskgpu_push_reads");
                bases.clear(); offs.assign(1, 0);
            }
        }
        if (offs.size() > 1) check(dskgpu_push_reads(ctx, bankId, bases.data(), offs.data(), offs.size() - 1), ctx, "dskgpu_push_reads");
    }

    void collectLeaves(IBank* b, std::vector<IBank*>& out)
    {
        const std::vector<IBank*> sub = b->getBanks();
        if (sub.size() == 1 && sub[0] == b) { out.push_back(b); return; }
        if (sub.empty()) { out.push_back(b); return; }
        for (size_t i = 0; i < sub.size(); i++) collectLeaves(sub[i], out);
    }

    /** Splits the banks into per-device task lists and runs one reader thread per device. */
    void feedBanks(int onlyBank = -1)
    {
        const size_t W = _ctxs.size();
        std::vector<std::vector<FeedTask> > tasks(W);
        size_t rr = 0;                                             // round robin for inputs that cannot be cut into ranges
        // bank id = index in the top-level composition, as the reference's per-bank counts (getCompositionNb)
        const std::vector<IBank*> top = _bank->getBanks();
        const bool composite = _config._nb_banks > 1 && top.size() == _config._nb_banks;
        for (size_t t = 0; t < (composite ? top.size() : 1); t++) {
            if (onlyBank >= 0 && (int)t != onlyBank) continue;     // plug-in mode with several banks: one device job per bank
            std::vector<IBank*> leaves;
            collectLeaves(composite ? top[t] : _bank, leaves);
            for (size_t i = 0; i < leaves.size(); i++) {
                const std::string id = leaves[i]->getId();
                FeedTask ft; ft.bank = (int)t; ft.path = id; ft.begin = 0; ft.end = 0; ft.whole = true; ft.leaf = leaves[i];
                if (_viaIterator || !isRegularFile(id)) { ft.path.clear(); tasks[rr++ % W].push_back(ft); continue; }
                const uint64_t size = (uint64_t)System::file().getSize(id);
                char first = 0;
                static const uint64_t splitMin = getenv("DSKGPU_SPLIT_MIN_BYTES") ? (uint64_t)atoll(getenv("DSKGPU_SPLIT_MIN_BYTES")) : ((uint64_t)8 << 20);
                if (W > 1 && size >= splitMin && !isGzip(id)) {
                    FILE* f = fopen(id.c_str(), "rb");
                    if (f) { int ch; while ((ch = fgetc(f)) != EOF) if (ch == '>' || ch == '@') { first = (char)ch; break; } fclose(f); }
                }
                if (!first) { tasks[rr++ % W].push_back(ft); continue; }       // gzip / small / unknown: one device parses it whole
                const int fd = open(id.c_str(), O_RDONLY);
                if (fd < 0) throw Exception("unable to open file %s", id.c_str());
                std::vector<uint64_t> cut(W + 1, size);
                cut[0] = 0;
                for (size_t r = 1; r < W; r++) cut[r] = std::max(cut[r - 1], nextRecordStart(fd, size / W * r, size, first));
                close(fd);
                for (size_t r = 0; r < W; r++) {
                    if (cut[r + 1] <= cut[r]) continue;
                    ft.whole = false; ft.begin = cut[r]; ft.end = cut[r + 1];
                    tasks[r].push_back(ft);
                }
            }
        }
        const size_t cap = (size_t)64 << 20;
        std::vector<ThreadError> errs(W);
        // whole files of a task list with at least two of them are inflated ahead by helper threads (DSKGPU_READ_AHEAD files
        // at a time, default 4 minus the devices sharing the host; 0 = the reader thread alone)
        const char* ra = getenv("DSKGPU_READ_AHEAD");
        const size_t ahead = ra ? (size_t)std::max(0, atoi(ra)) : (size_t)std::max<int>(1, 4 / (int)W);
        auto worker = [&](size_t r) {
            char* buf[2] = {(char*)dskgpu_host_alloc(cap), (char*)dskgpu_host_alloc(cap)};
            const std::vector<FeedTask>& tl = tasks[r];
            std::vector<size_t> files;                               // tasks served from a queue
            for (size_t i = 0; i < tl.size(); i++) if (!tl[i].path.empty() && tl[i].whole) files.push_back(i);
            const bool queued = ahead > 0 && files.size() >= 2;
            std::vector<BlockQueue*> queues(tl.size(), (BlockQueue*)0);
            std::vector<std::thread> helpers;
            size_t started = 0;                                      // files[0 .. started) have a helper
            auto startUpTo = [&](size_t n) {
                for (; started < std::min(n, files.size()); started++) {
                    BlockQueue* Q = new BlockQueue(); queues[files[started]] = Q;
                    helpers.push_back(std::thread(inflateFile, tl[files[started]].path, Q, (size_t)16 << 20, (size_t)4));
                }
            };
            try {
                if (!buf[0] || !buf[1]) throw Exception("pinned staging allocation failed");
                size_t served = 0;                                   // files already consumed
                for (size_t i = 0; i < tl.size(); i++) {
                    const FeedTask& ft = tl[i];
                    if (ft.path.empty()) feedSequences(_ctxs[r], ft.bank, ft.leaf);
                    else if (queued && ft.whole) { startUpTo(served + 1 + ahead); feedQueue(_ctxs[r], ft, queues[i]); served++; }
                    else feedRange(_ctxs[r], ft, buf, cap);
                }
            }
            catch (FormatRejected& e) { errs[r].failed = true; errs[r].format = true; errs[r].what = e.what; }
            catch (Exception& e) { errs[r].failed = true; errs[r].what = e.getMessage(); }
            catch (std::exception& e) { errs[r].failed = true; errs[r].what = e.what(); }
            for (size_t i = 0; i < queues.size(); i++) if (queues[i]) { std::unique_lock<std::mutex> lk(queues[i]->m); queues[i]->cancel = true; queues[i]->cv.notify_all(); }
            for (size_t i = 0; i < helpers.size(); i++) helpers[i].join();
            for (size_t i = 0; i < queues.size(); i++) delete queues[i];
            dskgpu_host_free(buf[0]); dskgpu_host_free(buf[1]);
        };
        if (W == 1) worker(0);
        else {
            std::vector<std::thread> th;
            for (size_t r = 0; r < W; r++) th.push_back(std::thread(worker, r));
            for (size_t r = 0; r < W; r++) th[r].join();
        }
        for (size_t r = 0; r < W; r++) if (errs[r].failed && errs[r].format) { FormatRejected e; e.what = errs[r].what; throw e; }
        for (size_t r = 0; r < W; r++) if (errs[r].failed) throw Exception("%s", errs[r].what.c_str());
    }

    /** exchange + counting on every device (dskgpu_finish for one, dskgpu_multi_finish for several) */
    void finishAll()
    {
        if (_ctxs.size() == 1) { checkFormat(dskgpu_finish(_ctx), _ctx, "dskgpu_finish"); return; }
        const int rc = dskgpu_multi_finish(_ctxs.data(), (int)_ctxs.size());
        if (rc != DSKGPU_OK) {
            dskgpu_ctx* bad = _ctx;
            for (size_t r = 0; r < _ctxs.size(); r++) { const char* d = dskgpu_last_error(_ctxs[r]); if (d && *d) { bad = _ctxs[r]; break; } }
            checkFormat(rc, bad, "dskgpu_multi_finish");
        }
    }

    void recountAll(const int64_t* amin)
    {
        const size_t W = _ctxs.size();
        std::vector<int> rcs(W, 0);
        if (W == 1) rcs[0] = dskgpu_recount(_ctx, amin);
        else {
            std::vector<std::thread> th;
            for (size_t r = 0; r < W; r++) th.push_back(std::thread([&, r] { rcs[r] = dskgpu_recount(_ctxs[r], amin); }));
            for (size_t r = 0; r < W; r++) th[r].join();
        }
        for (size_t r = 0; r < W; r++) check(rcs[r], _ctxs[r], "dskgpu_recount");
    }

    /** histograms of the job so far (this pass): summed over the devices */
    void sumHistograms(std::vector<uint64_t>& hs /*[nbHist][10001]*/, size_t nbHist, bool perBank)
    {
        std::vector<uint64_t> one(nbHist * (size_t)DSKGPU_HISTO_LEN);
        for (size_t r = 0; r < _ctxs.size(); r++) {
            if (perBank) check(dskgpu_bank_histograms(_ctxs[r], one.data()), _ctxs[r], "dskgpu_bank_histograms");
            else         check(dskgpu_histogram(_ctxs[r], one.data(), 0), _ctxs[r], "dskgpu_histogram");
            for (size_t i = 0; i < one.size(); i++) hs[i] += one[i];
        }
    }

    /** The pass loop of SortingCountAlgorithm::execute (K/SortingCountAlgorithm.cpp:678-689): every pass pushes the banks
     *  again and keeps its share of the minimizers.  "-abundance-min auto" needs the histogram of the WHOLE job before
     *  anything is dumped: with one pass the partitions are counted again from HBM (dskgpu_recount); with several, a first
     *  round of passes produces the histogram(s) and a second one dumps. */
    void runAllPasses()
    {
        const size_t W = _ctxs.size();
        const size_t nbHist = _autoPerBank ? (size_t)_config._nb_banks : 1;
        _h1.assign(DSKGPU_HISTO_LEN, 0); _h2.assign((size_t)DSKGPU_HISTO_LEN * DSKGPU_HISTO2D_DIM2, 0);
        memset(&_st, 0, sizeof _st); _nbSolidWritten = 0;
        Group& dsk = _storage->getGroup("dsk");
        _solidCounts = &dsk.template getPartition<Count>("solid", W * (size_t)_nbPasses);
        int64_t amin[DSKGPU_MAX_BANKS];
        bool haveThresholds = false;
        if (_autoCutoff && _nbPasses > 1) {
            std::vector<uint64_t> hs(nbHist * (size_t)DSKGPU_HISTO_LEN, 0);
            for (int pass = 0; pass < _nbPasses; pass++) {
                TIME_INFO(getTimeInfo(), "cutoff_passes");
                beginPass(pass);
                feedBanks();
                finishAll();
                sumHistograms(hs, nbHist, _autoPerBank);
            }
            thresholdsFrom(hs, nbHist, amin);
            haveThresholds = true;
        }
        for (int pass = 0; pass < _nbPasses; pass++) {
            if (_nbPasses > 1) beginPass(pass);
            {
                TIME_INFO(getTimeInfo(), "fill_partitions");          // same labels as K/SortingCountAlgorithm.cpp:1218
                feedBanks();
            }
            hostTrace("banks fed (file bytes queued to the device)");
            {
                TIME_INFO(getTimeInfo(), "fill_solid_kmers");         // K/SortingCountAlgorithm.cpp:1391
                finishAll();
                if (_autoCutoff) {
                    if (!haveThresholds) {
                        std::vector<uint64_t> hs(nbHist * (size_t)DSKGPU_HISTO_LEN, 0);
                        sumHistograms(hs, nbHist, _autoPerBank);
                        thresholdsFrom(hs, nbHist, amin);
                    }
                    recountAll(amin);
                }
            }
            hostTrace("finish done (counted, results on the host)");
            TIME_INFO(getTimeInfo(), "dump");
            collectPass(pass);
        }
        _solidCounts->flush();
    }

    // ---- plug-in mode: the loop of SortingCountAlgorithm::execute / fillSolidKmers (K/SortingCountAlgorithm.cpp:671-692,
    // :1391-1402, :1414-1607) with the device in the place of the partition commands ------------------------------------------
    void executeWithProcessors()
    {
        const size_t W = _ctxs.size();
        memset(&_st, 0, sizeof _st); _nbSolidWritten = 0;
        const size_t B = (size_t)std::max<size_t>(1, _config._nb_banks);
        if (B > 1) _config._nb_partitions = 1;                      // the merged stream of a pass is ONE ordered partition
        for (size_t i = 0; i < _processors.size(); i++) _processors[i]->begin(_config);
        // one device job: every distinct k-mer of `onlyBank` (or of everything) with its count
        auto runJob = [&](int pass, int onlyBank, bool resetFirst) {
            for (int attempt = 0; ; attempt++) {
                try {
                    if (resetFirst || attempt > 0) beginPass(pass);
                    { TIME_INFO(getTimeInfo(), "fill_partitions"); feedBanks(onlyBank); }
                    { TIME_INFO(getTimeInfo(), "fill_solid_kmers"); finishAll(); }
                    break;
                } catch (FormatRejected& e) {
                    if (attempt > 0 || _viaIterator) throw Exception("dskgpu: input rejected by the record scanner (%s)", e.what.c_str());
                    _viaIterator = true;
                }
            }
            for (size_t r = 0; r < W; r++) {
                dskgpu_stats st;
                check(dskgpu_get_stats(_ctxs[r], &st), _ctxs[r], "dskgpu_get_stats");
                if (pass == 0) { _st.nb_sequences += st.nb_sequences; _st.nb_nucleotides += st.nb_nucleotides; _st.kmers_nb_valid += st.kmers_nb_valid; }
                _st.nb_superkmers += st.nb_superkmers; _st.superkmer_bytes += st.superkmer_bytes; _st.gpu_launches += st.gpu_launches;
                if (B == 1) _st.kmers_nb_distinct += st.kmers_nb_distinct;
            }
        };
        // feeds one ordered stream of (k-mer, CountVector, sum) to every processor, the protocol of fillSolidKmers_aux
        for (int pass = 0; pass < _nbPasses; pass++) {
            if (B == 1) {
                runJob(pass, -1, _nbPasses > 1);
                TIME_INFO(getTimeInfo(), "processors");
                for (size_t i = 0; i < _processors.size(); i++) {
                    CountProcessor* proc = _processors[i];
                    proc->beginPass((size_t)pass);
                    std::vector<CountProcessor*> clones;
                    for (size_t r = 0; r < W; r++) {                   // one "partition" per device: its k-mers, ascending
                        CountProcessor* clone = proc->clone();
                        clone->use();
                        clones.push_back(clone);
                        const uint64_t* kmers = 0; const uint32_t* counts = 0; uint64_t n = 0; int words = 0;
                        check(dskgpu_partition(_ctxs[r], 0, &kmers, &counts, &n, &words), _ctxs[r], "dskgpu_partition");
                        clone->beginPart((size_t)pass, r, 200 * 1000, "device");
                        CountVector cv(1);
                        for (uint64_t j = 0; j < n; j++) {
                            Type v; setValue(v, kmers + j * words, words);
                            cv[0] = (CountNumber)counts[j];
                            clone->process(r, v, cv);                  // (sum left at 0, as PartitionsCommand.cpp:119,167 does: a chain computes it)
                        }
                        clone->endPart((size_t)pass, r);
                    }
                    proc->finishClones(clones);
                    for (size_t r = 0; r < clones.size(); r++) clones[r]->forget();
                    proc->endPass((size_t)pass);
                }
                continue;
            }
            // several banks: a CountVector holds one count per bank (K/PartitionsCommand.cpp:540-541).  Each bank is counted by a
            // device job of its own -- the counts of a k-mer in bank b are exactly what a job fed with bank b alone delivers --
            // and the W x B ordered lists are merged on the host (a k-mer's pass is a function of its minimizer, so with several
            // passes every list of a pass holds the same slice of the k-mer space).
            struct List { std::vector<uint64_t> keys; std::vector<uint32_t> cnt; size_t bank, pos; };
            std::vector<List> lists;
            int words = 1;
            for (size_t b = 0; b < B; b++) {
                runJob(pass, (int)b, _nbPasses > 1 || b > 0);
                for (size_t r = 0; r < W; r++) {
                    const uint64_t* kmers = 0; const uint32_t* counts = 0; uint64_t n = 0;
                    check(dskgpu_partition(_ctxs[r], 0, &kmers, &counts, &n, &words), _ctxs[r], "dskgpu_partition");
                    lists.push_back(List());
                    List& L = lists.back();
                    L.bank = b; L.pos = 0;
                    L.keys.assign(kmers, kmers + n * (uint64_t)words); L.cnt.assign(counts, counts + n);
                }
            }
            TIME_INFO(getTimeInfo(), "processors");
            const int KWn = words;
            auto less = [&](const uint64_t* a, const uint64_t* c) { for (int q = KWn - 1; q >= 0; q--) if (a[q] != c[q]) return a[q] < c[q]; return false; };
            auto same = [&](const uint64_t* a, const uint64_t* c) { for (int q = 0; q < KWn; q++) if (a[q] != c[q]) return false; return true; };
            // merged stream, materialised once (every processor walks it): keys + B counts per k-mer
            std::vector<uint64_t> mk; std::vector<CountNumber> mc;
            {
                std::vector<size_t> heap;                              // indices of the lists that still have items, min-heap on their head key
                auto head = [&](size_t l) { return lists[l].keys.data() + lists[l].pos * (size_t)KWn; };
                auto cmp = [&](size_t x, size_t y) { return less(head(y), head(x)); };   // std heap = max-heap: invert
                for (size_t l = 0; l < lists.size(); l++) if (!lists[l].cnt.empty()) heap.push_back(l);
                std::make_heap(heap.begin(), heap.end(), cmp);
                while (!heap.empty()) {
                    std::pop_heap(heap.begin(), heap.end(), cmp);
                    const size_t l = heap.back();
                    const uint64_t* k0 = head(l);
                    const bool fresh = mk.empty() || !same(&mk[mk.size() - (size_t)KWn], k0);
                    if (fresh) { mk.insert(mk.end(), k0, k0 + KWn); mc.insert(mc.end(), B, (CountNumber)0); }
                    mc[mc.size() - B + lists[l].bank] += (CountNumber)lists[l].cnt[lists[l].pos];
                    if (++lists[l].pos < lists[l].cnt.size()) std::push_heap(heap.begin(), heap.end(), cmp); else heap.pop_back();
                }
            }
            lists.clear();
            const uint64_t n = mk.size() / (size_t)KWn;
            _st.kmers_nb_distinct += n;
            for (size_t i = 0; i < _processors.size(); i++) {
                CountProcessor* proc = _processors[i];
                proc->beginPass((size_t)pass);
                std::vector<CountProcessor*> clones;
                CountProcessor* clone = proc->clone();
                clone->use();
                clones.push_back(clone);
                clone->beginPart((size_t)pass, 0, 200 * 1000, "device");
                CountVector cv(B);
                for (uint64_t j = 0; j < n; j++) {
                    Type v; setValue(v, &mk[j * (size_t)KWn], KWn);
                    for (size_t b = 0; b < B; b++) cv[b] = mc[j * B + b];
                    clone->process(0, v, cv);                          // (sum = 0: CountProcessorChain::computeSum applies -solidity-custom's vector)
                }
                clone->endPart((size_t)pass, 0);
                proc->finishClones(clones);
                clone->forget();
                proc->endPass((size_t)pass);
            }
        }
        for (size_t i = 0; i < _processors.size(); i++) _processors[i]->end();
        // statistics: the keys of K/SortingCountAlgorithm.cpp:728-780 that exist without the default chain
        getInfo()->add(1, "bank");
        getInfo()->add(2, "bank_uri", "%s", _bank->getId().c_str());
        getInfo()->add(2, "bank_total_nt", "%lld", (long long)_st.nb_nucleotides);
        getInfo()->add(2, "sequences");
        getInfo()->add(3, "seq_number", "%ld", (long)_st.nb_sequences);
        getInfo()->add(2, "kmers");
        getInfo()->add(3, "kmers_nb_valid", "%lld", (long long)_st.kmers_nb_valid);
        getInfo()->add(1, "stats");
        getInfo()->add(2, "temp_files");
        getInfo()->add(3, "nb_superkmers", "%lld", (long long)_st.nb_superkmers);
        getInfo()->add(2, "kmers");
        getInfo()->add(3, "kmers_nb_distinct", "%ld", (long)_st.kmers_nb_distinct);
        if (_processors.size() == 1) getInfo()->add(2, _processors[0]->getProperties());
        else for (size_t i = 0; i < _processors.size(); i++) { getInfo()->add(2, _processors[i]->getName()); getInfo()->add(3, _processors[i]->getProperties()); }
        getInfo()->add(1, getTimeInfo().getProperties("time"));
    }

    void beginPass(int pass)
    {
        for (size_t r = 0; r < _ctxs.size(); r++) {
            check(dskgpu_reset(_ctxs[r]), _ctxs[r], "dskgpu_reset");
            check(dskgpu_set_pass(_ctxs[r], pass, _nbPasses), _ctxs[r], "dskgpu_set_pass");
        }
    }

    // ---- -abundance-min auto: CountProcessorCutoff::endPass + CountProcessorSolidityInfo::setAbundanceMin ----------------
    // (K/CountProcessorCutoff.hpp:86-124, K/CountProcessorSolidity.hpp:45-66); the heuristic itself is the reference's own
    // Histogram::compute_threshold, run on the device histogram(s)
    void thresholdsFrom(const std::vector<uint64_t>& hs, size_t nbHist, int64_t* amin /*[DSKGPU_MAX_BANKS]*/)
    {
        _cutoffs.clear();
        for (size_t b = 0; b < nbHist; b++) {
            Histogram H(_histoMax);
            for (size_t i = 0; i <= _histoMax; i++) H.get((u_int16_t)i) = i < _histoMax ? hs[b * DSKGPU_HISTO_LEN + i] : 0;
            H.compute_threshold(3);
            _cutoffs.push_back((CountNumber)H.get_solid_cutoff());
        }
        const size_t nbThr = std::max<size_t>(1, _config._abundance.size());
        if (_cutoffs.size() > nbThr) throw Exception("Unable to set abundance min values (%d values for %d banks)", (int)_cutoffs.size(), (int)nbThr);
        for (size_t i = 0; i < _cutoffs.size(); i++) amin[i] = _userAbundanceMin[i] == -1 ? (int64_t)_cutoffs[i] : (int64_t)_userAbundanceMin[i];
        for (size_t i = _cutoffs.size(); i < (size_t)DSKGPU_MAX_BANKS; i++) amin[i] = amin[_cutoffs.size() - 1];
    }

    // ---- results of one pass: CountProcessorDump role (K/CountProcessorDump.hpp:85-152), in bulk -----------------------
    void collectPass(int pass)
    {
        const size_t W = _ctxs.size();
        std::vector<Count> block;
        std::vector<uint64_t> h1(DSKGPU_HISTO_LEN), h2((size_t)DSKGPU_HISTO_LEN * DSKGPU_HISTO2D_DIM2);
        for (size_t r = 0; r < W; r++) {
            dskgpu_ctx* ctx = _ctxs[r];
            const size_t coll = (size_t)pass * W + r;               // CountProcessorDump: partId + passId * nbPartsPerPass
            const uint64_t* kmers = 0; const uint32_t* counts = 0; uint64_t n = 0; int words = 0;
            check(dskgpu_partition(ctx, 0, &kmers, &counts, &n, &words), ctx, "dskgpu_partition");
            // The device hands back two arrays (keys, abundances) in pinned memory; the dataset is an array of Count{value,
            // abundance}.  Blocks of 4 M items are interleaved by a few threads (a straight word copy when Type is laid out
            // as its 64-bit words, least significant first -- checked on the first item -- else through the arithmetic
            // interface) and each block is ONE H5Dset_extent + H5Dwrite (BagHDF5::insert, CollectionHDF5.hpp:79-117): no
            // BagCache, no per-item insert, no synchronizer on the path.
            const size_t B = (size_t)4 << 20;
            const bool raw = n > 0 && rawLayoutOk(kmers, words);
            const size_t nthr = std::max<size_t>(1, std::min<size_t>(8, std::thread::hardware_concurrency()));
            for (uint64_t i = 0; i < n; i += B) {
                const size_t m = (size_t)std::min<uint64_t>(B, n - i);
                block.resize(m);
                auto fill = [&](size_t a, size_t b) {
                    for (size_t j = a; j < b; j++) {
                        if (raw) { memcpy((void*)&block[j].value, kmers + (i + j) * words, sizeof(uint64_t) * (size_t)words); block[j].abundance = (CountNumber)counts[i + j]; }
                        else { Type v; setValue(v, kmers + (i + j) * words, words); block[j] = Count(v, (CountNumber)counts[i + j]); }
                    }
                };
                if (m < ((size_t)1 << 16) || nthr == 1) fill(0, m);
                else {
                    std::vector<std::thread> th;
                    for (size_t t = 0; t < nthr; t++) th.push_back(std::thread(fill, m * t / nthr, m * (t + 1) / nthr));
                    for (size_t t = 0; t < nthr; t++) th[t].join();
                }
                (*_solidCounts)[coll].insert(block.data(), m);
            }
            (*_solidCounts)[coll].flush();
            _nbSolidWritten += n;
            check(dskgpu_histogram(ctx, h1.data(), h2.data()), ctx, "dskgpu_histogram");
            for (size_t i = 0; i < h1.size(); i++) _h1[i] += h1[i];
            for (size_t i = 0; i < h2.size(); i++) _h2[i] += h2[i];
            dskgpu_stats st;
            check(dskgpu_get_stats(ctx, &st), ctx, "dskgpu_get_stats");
            if (pass == 0) {                                         // every pass parses the whole bank: report it once
                _st.nb_sequences += st.nb_sequences; _st.nb_nucleotides += st.nb_nucleotides; _st.kmers_nb_valid += st.kmers_nb_valid;
                // BankStats::operator+= over the devices' slices (K/BankKmers.hpp:184-194)
                if (st.seq_stats_sequences) {
                    _st.seq_len_min = _st.seq_stats_sequences ? std::min(_st.seq_len_min, st.seq_len_min) : st.seq_len_min;
                    _st.seq_len_max = std::max(_st.seq_len_max, st.seq_len_max);
                }
                _st.seq_stats_sequences += st.seq_stats_sequences; _st.seq_len_sum += st.seq_len_sum; _st.seq_len_sumsq += st.seq_len_sumsq;
                _st.kmers_nb_invalid += st.kmers_nb_invalid;
            }
            _st.nb_superkmers += st.nb_superkmers; _st.superkmer_bytes += st.superkmer_bytes;
            _st.kmers_nb_distinct += st.kmers_nb_distinct; _st.kmers_nb_solid += st.kmers_nb_solid;
            if (r == 0) _st.nb_partitions += st.nb_partitions;      // the plan is job-wide: every device reports the same number
            _st.nb_groups_hash += st.nb_groups_hash; _st.nb_groups_sort += st.nb_groups_sort; _st.gpu_launches += st.gpu_launches;
            _st.ms_parse = std::max(_st.ms_parse, st.ms_parse); _st.ms_superk = std::max(_st.ms_superk, st.ms_superk);
            _st.ms_partition = std::max(_st.ms_partition, st.ms_partition); _st.ms_count = std::max(_st.ms_count, st.ms_count);
            _st.ms_sort = std::max(_st.ms_sort, st.ms_sort);
        }
    }

    // ---- results: CountProcessorDump / CountProcessorHistogram roles, in bulk ---------------------------------------
    void writeResults()
    {
        IProperties* in = getInput();

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

        // dsk/solid/<p> were written pass by pass (collectPass); attributes as CountProcessorDump::begin leaves them
        Group& dsk = _storage->getGroup("dsk");
        const int nparts = (int)(_ctxs.size() * (size_t)_nbPasses);
        dsk.addProperty("kmer_size", Stringify::format("%d", (int)_config._kmerSize));
        const u_int64_t nbSolid = _nbSolidWritten;

        // histogram group + text files: the reference's own CountProcessorHistogram::end() on a Histogram object that
        // was filled from the device bins (K/CountProcessorHistogram.hpp:104-159, Histogram.cpp:43-190)
        const std::vector<uint64_t>& h1 = _h1; const std::vector<uint64_t>& h2 = _h2;
        const bool histo2D = !_histo2DName.empty(), histo1D = !_histoName.empty();
        CountProcessorHistogram<span> ph(&_storage->getGroup("histogram"), _histoMax, in->getInt(STR_KMER_ABUNDANCE_MIN_THRESHOLD),
                                         histo2D, histo1D, _histo2DName, _histoName);
        IHistogram* H = ph.getHistogram();
        // 1-D: bins below the length as they are, the clamp bin stays empty (never merged: Histogram.hpp:221); 2-D: the clamp
        // column collects everything at or above the length (Histogram.hpp:97)
        for (size_t i = 0; i <= _histoMax; i++) H->get((u_int16_t)i) = i < _histoMax ? h1[i] : 0;
        for (size_t j = 0; j <= 10; j++) {
            for (size_t i = 0; i < _histoMax; i++) H->get2D((u_int16_t)i, (u_int16_t)j) = h2[j * DSKGPU_HISTO_LEN + i];
            u_int64_t tail = 0;
            for (size_t i = _histoMax; i <= 10000; i++) tail += h2[j * DSKGPU_HISTO_LEN + i];
            H->get2D((u_int16_t)_histoMax, (u_int16_t)j) = tail;
        }
        ph.end();

        // ---- statistics: keys of K/SortingCountAlgorithm.cpp:728-780 ----------------------------------------------------
        getInfo()->add(1, "bank");
        getInfo()->add(2, "bank_uri", "%s", _bank->getId().c_str());
        getInfo()->add(2, "bank_size", "%lld", (long long)_bank->getSize());
        getInfo()->add(2, "bank_total_nt", "%lld", (long long)_st.nb_nucleotides);
        getInfo()->add(2, "sequences");
        getInfo()->add(3, "seq_number", "%ld", (long)_st.nb_sequences);
        const double seqMean = _st.nb_sequences ? (double)_st.nb_nucleotides / (double)_st.nb_sequences : 0.0;
        if (_st.seq_stats_sequences) {                              // K/SortingCountAlgorithm.cpp:735-738, BankStats::getSeqDeviation
            getInfo()->add(3, "seq_size_min", "%ld", (long)_st.seq_len_min);
            getInfo()->add(3, "seq_size_max", "%ld", (long)_st.seq_len_max);
        }
        getInfo()->add(3, "seq_size_mean", "%.1f", seqMean);
        if (_st.seq_stats_sequences)
            getInfo()->add(3, "seq_size_deviation", "%.1f", sqrt(std::max(0.0, (double)_st.seq_len_sumsq / (double)_st.seq_stats_sequences - seqMean * seqMean)));
        getInfo()->add(2, "kmers");
        getInfo()->add(3, "kmers_nb_valid", "%lld", (long long)_st.kmers_nb_valid);
        if (_st.seq_stats_sequences) getInfo()->add(3, "kmers_nb_invalid", "%lld", (long long)_st.kmers_nb_invalid);
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
        getInfo()->add(3, "nb_passes", "%ld", (long)_nbPasses);
        getInfo()->add(3, "nb_devices", "%ld", (long)_ctxs.size());
        getInfo()->add(3, "minimizer_size_used", "%ld", (long)_minimizerSizeUsed);
        if (_minimizerType != 0 || _repartitionType != 0)
            getInfo()->add(3, "partition_balance", "%s", "-minimizer-type / -repartition-type accepted; the device path balances from exact minimizer-bin counts");
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
    /** true when a Type is stored as its `words` 64-bit words, least significant first (LargeInt<1>: one u64; LargeInt<2>:
     *  __uint128_t or u64[2]; LargeInt<3|4>: u64[N]) -- verified against the arithmetic interface on a real key */
    static bool rawLayoutOk(const uint64_t* w, int words)
    {
        if (sizeof(Type) != sizeof(uint64_t) * (size_t)words) return false;
        Type v; setValue(v, w, words);
        return memcmp((const void*)&v, w, sizeof(Type)) == 0;
    }

    static void setValue(Type& v, const uint64_t* w, int words)
    {
        v.setVal(w[words - 1]);
        for (int i = words - 2; i >= 0; i--) { v <<= 32; v <<= 32; Type lo; lo.setVal(w[i]); v = v + lo; }
    }
};

}  // namespace dskgpu_host
