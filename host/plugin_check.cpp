// plugin_check -- exercises the plug-in surface of GpuSortingCount<span> the way G/src/gatb/debruijn/impl/Graph.cpp:360-407
// uses SortingCountAlgorithm<span>: ConfigurationAlgorithm first, then the 5-argument constructor with the REFERENCE'S OWN
// count processors (SortingCountAlgorithm<span>::getDefaultProcessorVector: histogram -> solidity -> dump, preceded by the
// cutoff processor for "-abundance-min auto") plus one processor of our own that audits what the device path delivers:
// every distinct k-mer exactly once per partition, ascending, with its count.  The .h5 the reference's dump processor writes
// is then read back by the reference's dsk2ascii (tests/test_cli_dropin.py::test_plugin_surface_*).
//
//   plugin_check -file reads.fa -kmer-size 31 -abundance-min 2 -out prefix
#include "GpuSortingCount.hpp"

#include <gatb/kmer/impl/CountProcessorAbstract.hpp>

using namespace std;
using namespace dskgpu_host;

/** Audits the stream of (k-mer, counts) the processors are fed with. */
template <size_t span>
class AuditProcessor : public CountProcessorAbstract<span>
{
public:
    typedef typename Kmer<span>::Type Type;
    struct Totals { u_int64_t distinct, occurrences, unordered, badVector, parts; Totals() : distinct(0), occurrences(0), unordered(0), badVector(0), parts(0) {} };

    AuditProcessor(size_t nbBanks, Totals* shared = 0) : CountProcessorAbstract<span>("audit"), _nbBanks(nbBanks), _shared(shared ? shared : &_own), _first(true) {}

    CountProcessorAbstract<span>* clone() { return new AuditProcessor(_nbBanks, _shared); }
    void beginPart(size_t passId, size_t partId, size_t cacheSize, const char* name) { _first = true; _local = Totals(); }
    bool process(size_t partId, const Type& kmer, const CountVector& count, CountNumber sum)
    {
        // one count per bank, at least one occurrence, sum left to the processors (the reference passes 0: PartitionsCommand.cpp:119)
        CountNumber total = 0; bool neg = false;
        for (size_t b = 0; b < count.size(); b++) { total += count[b]; neg = neg || count[b] < 0; }
        if (count.size() != _nbBanks || neg || total < 1 || sum != 0) _local.badVector++;
        if (!_first && !(_prev < kmer)) _local.unordered++;         // ascending inside a partition (K/PartitionsCommand.cpp:540-541)
        _prev = kmer; _first = false;
        _local.distinct++; _local.occurrences += (u_int64_t)total;
        return true;
    }
    void endPart(size_t passId, size_t partId)
    {
        __sync_fetch_and_add(&_shared->distinct, _local.distinct); __sync_fetch_and_add(&_shared->occurrences, _local.occurrences);
        __sync_fetch_and_add(&_shared->unordered, _local.unordered); __sync_fetch_and_add(&_shared->badVector, _local.badVector);
        __sync_fetch_and_add(&_shared->parts, 1);
    }
    const Totals& totals() const { return *_shared; }

private:
    size_t _nbBanks; Totals _own, _local; Totals* _shared;
    Type _prev; bool _first;
};

struct Parameter { IProperties* props; };

template <size_t span> struct Functor { void operator()(Parameter parameter)
{
    IProperties* props = parameter.props;
    IBank* bank = Bank::open(props->getStr(STR_URI_FILE));
    LOCAL(bank);
    // what Graph.cpp does before SortingCountAlgorithm: storage, configuration, default processors
    Storage* storage = StorageFactory(STORAGE_HDF5).create(props->getStr(STR_URI_OUTPUT), true, false);
    LOCAL(storage);
    ConfigurationAlgorithm<span> configAlgo(bank, props);
    configAlgo.execute();
    Configuration config = configAlgo.getConfiguration();
    std::vector<ICountProcessor<span>*> processors = SortingCountAlgorithm<span>::getDefaultProcessorVector(config, props, storage, storage);
    AuditProcessor<span>* audit = new AuditProcessor<span>(config._nb_banks);
    processors.push_back(audit);

    GpuSortingCount<span> sortingCount(bank, config, 0, processors, props);
    sortingCount.execute();

    const typename AuditProcessor<span>::Totals& t = audit->totals();
    cout << "audit distinct " << t.distinct << " occurrences " << t.occurrences << " unordered " << t.unordered
         << " bad_vectors " << t.badVector << " parts " << t.parts << " processors " << sortingCount.getProcessorNumber() << endl;
} };

int main(int argc, char* argv[])
{
    try {
        IOptionsParser* parser = SortingCountAlgorithm<>::getOptionsParser();
        LOCAL(parser);
        if (IOptionsParser* input = parser->getParser(STR_URI_INPUT)) { input->setName(STR_URI_FILE); }
        IProperties* props = parser->parse(argc, argv);
        Parameter p; p.props = props;
        Integer::apply<Functor, Parameter>(props->getInt(STR_KMER_SIZE), p);
    }
    catch (OptionFailure& e) { return e.displayErrors(std::cout); }
    catch (Exception& e) { cerr << "EXCEPTION: " << e.getMessage() << endl; return EXIT_FAILURE; }
    return EXIT_SUCCESS;
}
