// Headless stand-in for <wrl/client.h> (oracle build only).  Nothing on the CPU
// simulation path instantiates ComPtr.
#pragma once
