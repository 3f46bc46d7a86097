// TEST INFRASTRUCTURE ONLY: src/relocator.cpp includes it, nothing of it is used on the paths compiled here.
#pragma once
