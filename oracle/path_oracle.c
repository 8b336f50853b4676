/* placeholder, replaced below */
int g19o_path_placeholder(void) { return 0; }
