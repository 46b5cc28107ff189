// dsk_gpu -- the `dsk` command line (R/src/main.cpp:28-49 + R/src/DSK.cpp:45-104) with the counting class replaced by
// GpuSortingCount<span> (libdskgpu.so behind include/dskgpu.h).  Same Tool framework, same option parser, same
// Integer::apply dispatch over KSIZE_LIST, same statistics/XML handling -- all of it gatb-core library code.
#include "GpuSortingCount.hpp"

using namespace std;
using namespace dskgpu_host;

class DSKGpu : public Tool
{
public:
    DSKGpu() : Tool("dsk")
    {
        getParser()->push_back(GpuSortingCount<>::getOptionsParser(), 1);
        if (IOptionsParser* input = getParser()->getParser(STR_URI_INPUT)) { input->setName(STR_URI_FILE); }
    }
    void execute();
};

struct Parameter
{
    Parameter(DSKGpu& dsk, IProperties* props) : dsk(dsk), props(props) {}
    DSKGpu&      dsk;
    IProperties* props;
};

template <size_t span> struct Functor { void operator()(Parameter parameter)
{
    DSKGpu&      dsk   = parameter.dsk;
    IProperties* props = parameter.props;

    IBank* bank = Bank::open(props->getStr(STR_URI_FILE));
    LOCAL(bank);

    GpuSortingCount<span> sortingCount(bank, props);
    sortingCount.keepMinimizerSize(dsk.getParser()->saw(STR_MINIMIZER_SIZE));     // an explicit -minimizer-size is honoured as given
    sortingCount.getInput()->add(0, STR_VERBOSE, props->getStr(STR_VERBOSE));
    sortingCount.execute();

    dsk.getInfo()->add(1, sortingCount.getConfig().getProperties());
    dsk.getInfo()->add(1, sortingCount.getInfo());
    sortingCount.getStorage()->getGroup(sortingCount.getName()).setProperty("xml", string("\n") + sortingCount.getInfo()->getXML());
} };

void DSKGpu::execute()
{
    size_t kmerSize = getInput()->getInt(STR_KMER_SIZE);
    Integer::apply<Functor, Parameter>(kmerSize, Parameter(*this, getInput()));
}

int main(int argc, char* argv[])
{
    try
    {
        hostTrace("main");
        DSKGpu().run(argc, argv);
        hostTrace("run returned");
    }
    catch (OptionFailure& e)
    {
        return e.displayErrors(std::cout);
    }
    catch (Exception& e)
    {
        cerr << "EXCEPTION: " << e.getMessage() << endl;
        return EXIT_FAILURE;
    }
    return EXIT_SUCCESS;
}
