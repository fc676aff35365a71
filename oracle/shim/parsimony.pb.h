#pragma once
// oracle shim: the protobuf load/save bodies are cut out of the oracle build (see oracle/Makefile), so
// only the names need to exist.
namespace Parsimony { struct data {}; }
namespace google { namespace protobuf { inline void ShutdownProtobufLibrary() {} } }
#define GOOGLE_PROTOBUF_VERIFY_VERSION
