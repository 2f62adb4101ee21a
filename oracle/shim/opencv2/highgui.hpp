// Test-infrastructure shim (NOT product code): highgui is only used under DEBUG_CV.
