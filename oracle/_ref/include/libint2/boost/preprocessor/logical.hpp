# /* Copyright (C) 2001
#  * Housemarque Oy
#  * http://www.housemarque.com
#  *
#  * Distributed under the Boost Software License, Version 1.0. (See
#  * accompanying file LICENSE_1_0.txt or copy at
#  * http://www.boost.org/LICENSE_1_0.txt)
#  */
#
# /* Revised by Paul Mensonides (2002) */
#
# /* See http://www.boost.org for most recent version. */
#
# ifndef BOOST_PREPROCESSOR_LOGICAL_HPP
# define BOOST_PREPROCESSOR_LOGICAL_HPP
#
# include <libint2/boost/preprocessor/logical/and.hpp>
# include <libint2/boost/preprocessor/logical/bitand.hpp>
# include <libint2/boost/preprocessor/logical/bitnor.hpp>
# include <libint2/boost/preprocessor/logical/bitor.hpp>
# include <libint2/boost/preprocessor/logical/bitxor.hpp>
# include <libint2/boost/preprocessor/logical/bool.hpp>
# include <libint2/boost/preprocessor/logical/compl.hpp>
# include <libint2/boost/preprocessor/logical/nor.hpp>
# include <libint2/boost/preprocessor/logical/not.hpp>
# include <libint2/boost/preprocessor/logical/or.hpp>
# include <libint2/boost/preprocessor/logical/xor.hpp>
#
# endif
