// Headless stand-in for <DirectXColors.h> (oracle build only); the seeders include it
// but use no named colour.
#pragma once
